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

# ---- section breakdown of our step (synchronised sections: sums exceed the pipelined step time)
import collections
acc = collections.OrderedDict()
def sec(name, t0):
    torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
for i in range(10):
    xb = x[(i % 4) * B:(i % 4 + 1) * B]
    torch.cuda.synchronize(); t0 = time.perf_counter(); opt.zero_grad(); sec("zero_grad", t0)
    t0 = time.perf_counter(); z = m.encoder(xb); sec("encoder_fwd", t0)
    t0 = time.perf_counter(); xq, rq_loss, idx = m.rq(z, use_sk=True); sec("rq_fwd", t0)
    t0 = time.perf_counter(); out = m.decoder(xq); sec("decoder_fwd", t0)
    t0 = time.perf_counter(); loss, rec = m.compute_loss(out, rq_loss, xs=xb); sec("loss", t0)
    t0 = time.perf_counter(); loss.backward(); sec("backward", t0)
    t0 = time.perf_counter(); torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0); sec("clip", t0)
    t0 = time.perf_counter(); opt.step(); sec("optimizer", t0)
print(json.dumps({k: round(v / 10, 3) for k, v in acc.items()}))

# ---- GPU time of the native MLP forward / backward calls alone
from lcrec_b200 import ops
h = m.encoder._get_handle()
xb = x[:B]
y, acts = h.forward(xb, want_acts=True)
gy = torch.randn_like(y)
for name, fn in (("encoder_mlp_forward", lambda: h.forward(xb, want_acts=True)), ("encoder_mlp_backward", lambda: h.backward(xb, acts, gy, need_gx=False)),
                 ("encoder_weight_update", lambda: h.update(h._keep, h._keepb))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): fn()
    e1.record(); host_ms = (time.perf_counter() - t0) / 10 * 1e3; torch.cuda.synchronize()
    print(json.dumps({name + "_gpu_ms": round(e0.elapsed_time(e1) / 10, 3), "host_issue_ms": round(host_ms, 3)}))
