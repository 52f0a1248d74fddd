"""Timing of the residual quantiser: SIMT kernels vs the tensor-core distance path (standard and large-codebook shapes)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
z = torch.randn(1_000_000, 32, device=dev, generator=g)
cbs = [torch.randn(256, 32, device=dev, generator=g) * 0.7 * 0.6 ** l for l in range(4)]
ops.rq_set_tc_mode(0)
for ipt in (2, 4, 2, 4):
    ops.rq_set_tc_mode(10 + ipt)
    ms = t(lambda: ops.rq_quantize(z, cbs, resid_level=3))
    print(json.dumps({"simt_items_per_thread": ipt, "ms_per_1M": round(ms, 3)}), flush=True)
ops.rq_set_tc_mode(12); ops.rq_set_tc_mode(1)
for n, d, ks in ((1_000_000, 32, [256] * 4),):
    z = torch.randn(n, d, device=dev, generator=g)
    cbs = [torch.randn(k, d, device=dev, generator=g) * 0.7 * 0.6 ** l for l, k in enumerate(ks)]
    out = {"n": n, "e_dim": d, "codes": ks}
    for mode, name in ((0, "simt"), (2, "tensor_core")):
        ops.rq_set_tc_mode(mode)
        res = ops.rq_quantize(z, cbs, resid_level=len(ks) - 1)
        ms = t(lambda: ops.rq_quantize(z, cbs, resid_level=len(ks) - 1), it=3 if (mode == 0 and d == 256) else 5)
        out[name + "_ms"] = round(ms, 3)
        out[name + "_items_per_s"] = round(n / ms * 1e3)
        if mode == 0: ref = res["codes"]
        else:
            out["rows_differing"] = int((res["codes"] != ref).any(dim=1).sum())
            flop = 2.0 * n * d * sum(ks)
            out["tc_algorithmic_tflops"] = round(flop / ms / 1e9, 1)
    ops.rq_set_tc_mode(1)
    print(json.dumps(out), flush=True)
