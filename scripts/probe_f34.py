"""Timing of the 8(f) kernels on one B200 (CUDA events, warm): masked mean pool (HBM-bound: bytes of hidden state read /
time vs the measured copy peak) and the EMA codebook step.  Prints one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lcrec_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for b, t, h, dt in [(64, 512, 4096, torch.float32), (256, 512, 4096, torch.float16), (1, 2048, 4096, torch.float32),
                        (1024, 128, 4096, torch.bfloat16)]:
        x = torch.randn(b, t, h, device=DEV, dtype=dt)
        mask = torch.ones(b, t, dtype=torch.int64, device=DEV)
        out = torch.empty(b, h, device=DEV)
        ms = timed(lambda: ops.masked_mean_pool(x, mask, out=out))
        ref = timed(lambda: (x * mask.unsqueeze(-1)).sum(dim=1) / mask.sum(dim=-1, keepdim=True))
        gb = x.numel() * x.element_size() / 1e9
        print(json.dumps({"kernel": "masked_mean_pool", "shape": [b, t, h], "dtype": str(dt), "ms": ms, "GB/s": gb / ms * 1e3,
                          "torch_expr_ms": ref, "input_GB": gb}))
        del x
    for n, k, d in [(1024, 256, 32), (65536, 256, 32), (1 << 20, 256, 32), (1024, 8192, 256)]:
        lat = torch.randn(n, d, device=DEV)
        idx = torch.randint(0, k, (n,), device=DEV)
        cs, w, cb = torch.zeros(k, device=DEV), torch.zeros(k, d, device=DEV), torch.randn(k, d, device=DEV)
        ms = timed(lambda: ops.ema_update(lat, idx, cs, w, cb, 0.99, 1e-5))

        def torch_ema():
            c = torch.zeros(k, device=DEV).scatter_add_(0, idx, torch.ones(n, device=DEV))
            cs.mul_(0.99).add_(c, alpha=0.01)
            dw = torch.zeros(k, d, device=DEV).index_add_(0, idx, lat)
            w.mul_(0.99).add_(dw, alpha=0.01)
        ref = timed(torch_ema)
        print(json.dumps({"kernel": "ema_update", "n": n, "codes": k, "e_dim": d, "ms": ms, "torch_atomic_ms": ref}))
    del flush


if __name__ == "__main__":
    main()
