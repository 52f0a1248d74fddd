"""What bounds the layer-1 GEMM?  Times the f16 x3 kernels (CTA-pair and single-CTA) with parts of the pipeline
switched off (lcrec_mlp_set_variant bits: 4 = no TMA loads after the first ring fill, 8 = no MMAs, 16 = single-CTA
kernel instead of the pair kernel, 32 = pair kernel with 6 x 32 KB stages instead of 3 x 64 KB)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
n = 131072
x = torch.randn(n, 4096, device=dev, generator=g)
ws = [torch.randn(2048, 4096, device=dev, generator=g) * 0.02]
bs = [torch.zeros(2048, device=dev)]
ops.set_default_engine(1)
h = ops.MlpHandle(ws, bs)
ref = (x[:512].double() @ ws[0].double().t()).float()
names = {0: "full", 4: "no TMA (MMA + fold only)", 8: "no MMA (TMA ingest only)", 12: "neither (handshakes + fold + epilogue)"}
for rep in range(2):
    for kern, kname in ((0, "pair BK64x3"), (32, "pair BK32x6"), (16, "single BK32x4")):
        for dbg in (0, 4, 8, 12):
            h.set_variant(kern | dbg)
            for _ in range(3): y = h.forward(x)
            ops.profile_enable(True); ops.profile_collect()
            for _ in range(10): y = h.forward(x)
            prof = ops.profile_collect(); ops.profile_enable(False)
            ms = prof[1][0] / prof[1][1]
            err = float((y[:512] - ref).abs().max() / ref.abs().max()) if dbg == 0 else None
            print(json.dumps(dict(kernel=kname, what=names[dbg], gemm_ms=round(ms, 3), tflops=round(2.0 * n * 4096 * 2048 / ms / 1e9, 1),
                                  split_ms=round(prof[0][0] / prof[0][1], 3), rel_err=err)), flush=True)
