"""Is the pair GEMM power-bound?  Layer-1 GEMM time with the persistent grid capped to fewer CTA pairs."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops, _lib
import bench
dev = torch.device("cuda:0")
lib = _lib.load()
ws_, bs_, cbs, head = bench.make_model()
x = bench.make_items_device(131072, head, dev, 0)
ws = [torch.from_numpy(ws_[0]).to(dev)]; bs = [torch.from_numpy(bs_[0]).to(dev)]
ops.set_default_engine(1)
h = ops.MlpHandle(ws, bs)
for rep in range(2):
    for cap in (0, 72, 68, 64, 60, 56, 48):
        lib.lcrec_pair_set_cluster_cap(cap)
        for _ in range(3): h.forward(x)
        ops.profile_enable(True); ops.profile_collect()
        for _ in range(12): h.forward(x)
        prof = ops.profile_collect(); ops.profile_enable(False)
        ms = prof[1][0] / prof[1][1]
        print(json.dumps(dict(cap=cap or 74, gemm_ms=round(ms, 3), tflops=round(2.0 * 131072 * 4096 * 2048 / ms / 1e9, 1))), flush=True)
lib.lcrec_pair_set_cluster_cap(0)
