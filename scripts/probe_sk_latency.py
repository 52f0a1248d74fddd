"""Latency / throughput of the per-group Sinkhorn warp kernels: G groups of n rows, scaling mode (no literal re-run)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
cb = torch.randn(256, 32, device=dev, generator=g) * 0.1
for mode in (1, 2):
    ops.sinkhorn_set_mode(mode)
    for n in (2, 3, 8):
        for G in (1, 64, 2368, 20000):
            rows = G * n
            resid = torch.randn(rows, 32, device=dev, generator=g) * 0.1
            base = resid[::n].repeat_interleave(n, 0)
            resid = base + 1e-3 * torch.randn(rows, 32, device=dev, generator=g)
            off = torch.arange(0, rows + 1, n, device=dev, dtype=torch.int64)
            mem = torch.arange(rows, device=dev, dtype=torch.int64)
            codes = torch.zeros(rows, 4, dtype=torch.int64, device=dev)
            ng = torch.tensor([G], device=dev, dtype=torch.int64)
            for _ in range(3): ops.sinkhorn_groups(resid, cb, off, mem, ng, G, rows, 0.003, 50, codes, 3)
            ops.profile_enable(False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps): ops.sinkhorn_groups(resid, cb, off, mem, ng, G, rows, 0.003, 50, codes, 3)
            e1.record(); torch.cuda.synchronize()
            print(json.dumps(dict(mode=mode, rows_per_group=n, groups=G, us_per_call=round(e0.elapsed_time(e1) / reps * 1e3, 1))), flush=True)
ops.sinkhorn_set_mode(2)
