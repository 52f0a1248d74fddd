"""Training step of BASELINE.json configs[1] with the fused RQ autograd node + deferred host checks vs the per-level path
(same script as probe_train.py's first section, both variants back to back).  One JSON line."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
from lcrec_b200.models import RQVAE
from lcrec_b200.models import rq as RQ
dev = torch.device("cuda:0")
dims = [2048, 1024, 512, 256, 128, 64]
B, steps = 1024, 40
x = torch.randn(B * 4, 4096, device=dev)
out = {}
from lcrec_b200.optim import FusedAdam
for name, fused, defer, fopt in (("per_level", False, False, False), ("fused_deferred", True, True, False),
                                 ("fused_deferred_fusedopt", True, True, True)):
    torch.manual_seed(0)
    m = RQVAE(in_dim=4096, num_emb_list=[256] * 4, e_dim=32, layers=dims, sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50).to(dev).train()
    opt = (FusedAdam if fopt else torch.optim.AdamW)(m.parameters(), lr=1e-3, weight_decay=1e-4)
    RQ.FUSED_TRAIN = fused

    def step(i):
        xb = x[(i % 4) * B:(i % 4 + 1) * B]
        opt.zero_grad()
        o, rq_loss, idx = m(xb, use_sk=True)
        loss, rec = m.compute_loss(o, rq_loss, xs=xb)
        loss.backward()
        if fopt:
            opt.clip_and_step(1.0)
        else:
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
        return loss

    def run(i):
        if defer:
            with ops.defer_checks():
                l = step(i)
                return l.item()
        return step(i).item()
    for i in range(5): run(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps): last = run(i)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    out[name + "_ms_per_step"] = round(ms, 3); out[name + "_loss"] = last
RQ.FUSED_TRAIN = True
print(json.dumps(out))
