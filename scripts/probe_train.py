"""Training step of BASELINE.json configs[1] (batch 1024, 4 x 256 codes, e_dim 32, encoder 4096-...-32, AdamW, clip 1.0,
Sinkhorn on the last level): this repo's model (forward + backward on the split-operand tensor-core kernels) against the
same architecture in plain torch (cuBLAS fp32, argmin quantiser without Sinkhorn as a lower bound of its work)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200.models import RQVAE
dev = torch.device("cuda:0")
torch.manual_seed(0)
dims = [2048, 1024, 512, 256, 128, 64]
B, steps = 1024, 25
x = torch.randn(B * 4, 4096, device=dev)
m = RQVAE(in_dim=4096, num_emb_list=[256] * 4, e_dim=32, layers=dims, sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50).to(dev).train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)

def step_ours(i):
    xb = x[(i % 4) * B:(i % 4 + 1) * B]
    opt.zero_grad()
    out, rq_loss, idx = m(xb, use_sk=True)
    loss, rec = m.compute_loss(out, rq_loss, xs=xb)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    opt.step()
    return loss

class TorchRef(torch.nn.Module):
    def __init__(self):
        super().__init__()
        d = [4096] + dims + [32]
        enc, dec = [], []
        for a, b in zip(d[:-1], d[1:]): enc += [torch.nn.Linear(a, b), torch.nn.ReLU()]
        for a, b in zip(d[::-1][:-1], d[::-1][1:]): dec += [torch.nn.Linear(a, b), torch.nn.ReLU()]
        self.enc, self.dec = torch.nn.Sequential(*enc[:-1]), torch.nn.Sequential(*dec[:-1])
        self.cbs = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(256, 32) * 0.01) for _ in range(4)])
    def forward(self, xb):
        z = self.enc(xb); r, xq, losses = z, 0, []
        for cb in self.cbs:
            dd = (r.detach() ** 2).sum(1, keepdim=True) + (cb ** 2).sum(1)[None] - 2 * r.detach() @ cb.t()
            q = cb[dd.argmin(1)]
            losses.append(torch.nn.functional.mse_loss(q, r.detach()) + 0.25 * torch.nn.functional.mse_loss(q.detach(), r))
            xr = r + (q - r).detach(); r = r - xr; xq = xq + xr
        return torch.nn.functional.mse_loss(self.dec(xq), xb) + torch.stack(losses).mean()
ref = TorchRef().to(dev).train()
opt2 = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-4)
def step_ref(i):
    xb = x[(i % 4) * B:(i % 4 + 1) * B]
    opt2.zero_grad(); loss = ref(xb); loss.backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0); opt2.step()
    return loss
out = {}
for name, fn in (("ours", step_ours), ("torch_fp32", step_ref)):
    for i in range(5): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps): fn(i)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    out[name + "_ms_per_step"] = round(ms, 3); out[name + "_items_per_s"] = round(B / ms * 1e3)
print(json.dumps(out))
