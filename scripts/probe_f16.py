"""Accuracy + speed of the f16 x3 engine vs the tf32 x3 engine vs torch fp32 (GPU box)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False

def time_fn(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def case(n, k, m, kind, chunk, timing=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, k, device=dev, generator=g)
    if kind == "relu":       # like hidden activations: half zeros, positive
        x = x.clamp_min(0)
    elif kind == "wide":     # 9 decades of dynamic range inside every row
        x = x * torch.exp(torch.empty(n, k, device=dev).uniform_(-10, 10, generator=g))
    elif kind == "rowscale": # rows of very different magnitude
        x = x * torch.exp(torch.empty(n, 1, device=dev).uniform_(-20, 20, generator=g))
    w = torch.randn(m, k, device=dev, generator=g) * (2.0 / (k + m)) ** 0.5
    b = torch.randn(m, device=dev, generator=g) * 0.01
    ref = x.double() @ w.double().t() + b.double()
    scale = (x.double().abs() @ w.double().abs().t()).mean().item() + 1e-300
    rowscale = (x.double().abs() @ w.double().abs().t()).mean(dim=1, keepdim=True) + 1e-300
    out = dict(n=n, k=k, m=m, kind=kind, chunk=chunk)
    for name, var in (("tf32", 0), ("f16", 2)):
        y = ops.linear_forward(x, w, b, False, acc_chunk=chunk, variant=var)
        err = (y.double() - ref)
        out[name + "_rel"] = float((err.abs() / rowscale).mean().item())
        out[name + "_max"] = float((err.abs() / rowscale).max().item())
        if timing:
            ms = time_fn(lambda: ops.linear_forward(x, w, b, False, acc_chunk=chunk, variant=var))
            out[name + "_ms"] = ms
            out[name + "_tflops"] = 2.0 * n * k * m / ms / 1e9
    t32 = torch.nn.functional.linear(x, w, b)
    out["torch_rel"] = float(((t32.double() - ref).abs() / rowscale).mean().item())
    print(json.dumps(out), flush=True)

for (n, k, m) in [(128, 32, 32), (300, 100, 48), (1000, 512, 512), (640, 16, 96)]:
    case(n, k, m, "normal", 64)
for kind in ("normal", "relu", "wide", "rowscale"):
    for chunk in (0, 64, 128, 256):
        case(2048, 4096, 2048, kind, chunk)
case(65536, 4096, 2048, "normal", 64, timing=True)
case(65536, 4096, 2048, "normal", 128, timing=True)
case(65536, 2048, 1024, "relu", 64, timing=True)
