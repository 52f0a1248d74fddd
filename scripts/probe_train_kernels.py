"""Kernel-level time of one training step of BASELINE configs[1] (batch 1024, run.sh architecture): meant to run under
`ncu --metrics gpu__time_duration.sum --clock-control none --csv` (launch list) or plain (wall time per step)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
from lcrec_b200.models import RQVAE
from lcrec_b200.optim import FusedAdam
dev = torch.device("cuda:0")
bn = "--bn" in sys.argv
steps = 3 if "--short" in sys.argv else 40
dims = [2048, 1024, 512, 256, 128, 64]
B = 1024
x = torch.randn(B * 4, 4096, device=dev)
torch.manual_seed(0)
m = RQVAE(in_dim=4096, num_emb_list=[256] * 4, e_dim=32, layers=dims, bn=bn, sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50).to(dev).train()
opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)


def step(i):
    xb = x[(i % 4) * B:(i % 4 + 1) * B]
    with ops.defer_checks():
        opt.zero_grad()
        o, rq_loss, idx = m(xb, use_sk=True)
        loss, rec = m.compute_loss(o, rq_loss, xs=xb)
        loss.backward()
        opt.clip_and_step(1.0)
        return loss.item()


for i in range(3):
    step(i)
torch.cuda.synchronize()
n0 = ops.launch_count()
torch.cuda.profiler.start()
t0 = time.perf_counter()
for i in range(steps):
    last = step(i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"bn": bn, "ms_per_step": round((time.perf_counter() - t0) / steps * 1e3, 3), "loss": last,
                  "own_launches_per_step": (ops.launch_count() - n0) / steps}))
