import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
dev = torch.device("cuda:0")
def time_fn(fn, iters=6, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
g = torch.Generator(device=dev).manual_seed(0)
n = 131072
x = torch.randn(n, 4096, device=dev, generator=g)
ws = [torch.randn(2048, 4096, device=dev, generator=g) * 0.02]
bs = [torch.zeros(2048, device=dev)]
ref = None
for eng in (1, 0):
    ops.set_default_engine(eng)
    h = ops.MlpHandle(ws, bs)
    for variant in (0, 1):
        h.set_variant(variant)
        for chunk in (0, 32, 64, 128, 256, 512, 1024):
            h.set_acc_chunk(chunk)
            ops.profile_enable(True); ops.profile_collect()
            for _ in range(4): y = h.forward(x)
            prof = ops.profile_collect(); ops.profile_enable(False)
            ms = prof[1][0] / prof[1][1]
            print(json.dumps(dict(engine="f16" if eng else "tf32", cluster=(variant == 0), chunk=chunk, gemm_ms=round(ms, 3),
                                  tflops=round(2.0 * n * 4096 * 2048 / ms / 1e9, 1), split_ms=round(prof[0][0] / prof[0][1], 3))), flush=True)
