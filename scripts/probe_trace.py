"""Pipeline timeline of one CTA of the layer-1 pair GEMM (clock64 stamps, see lcrec_mlp_set_trace)."""
import json, os, sys, torch, numpy as np
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
for variant in (0, 12):
    h.set_variant(variant)
    for _ in range(2): h.forward(x)
    tr = torch.zeros(6 * 512 * 4, dtype=torch.int64, device=dev)
    h.set_trace(tr); h.forward(x); torch.cuda.synchronize(); h.set_trace(None)
    t = tr.cpu().numpy().reshape(6, 512, 4)
    out = {"variant": variant}
    prod, mma, chunk, fold = t[0], t[1], t[2], t[3]
    sl = slice(70, 400)          # steady state (second tile onwards)
    out["producer_wait_empty_mean"] = float((prod[sl, 1] - prod[sl, 0]).mean())
    out["mma_wait_full_mean"] = float((mma[sl, 1] - mma[sl, 0]).mean())
    out["mma_issue_to_commit_mean"] = float((mma[sl, 2] - mma[sl, 1]).mean())
    out["kblock_period_mean"] = float(np.diff(mma[sl, 1]).mean())
    csl = slice(40, 200)
    out["mma_wait_tempty_mean"] = float((chunk[csl, 1] - chunk[csl, 0]).mean())
    out["fold_wait_tfull_mean"] = float((fold[csl, 1] - fold[csl, 0]).mean())
    out["fold_work_mean"] = float((fold[csl, 2] - fold[csl, 1]).mean())
    out["chunk_period_mean"] = float(np.diff(fold[csl, 1]).mean())
    # latency from "MMA thread commits empty(s) for k-block i" to "producer sees empty for k-block i + STAGES"
    stages = 6 if variant & 32 else 3
    lat = prod[70 + stages:400 + stages, 1] - mma[70:400, 2]
    out["commit_to_producer_wakeup_mean"] = float(lat.mean()); out["commit_to_producer_wakeup_min"] = float(lat.min())
    # latency from producer issue (after wait) to MMA thread seeing full
    lat2 = mma[sl, 1] - prod[sl, 1]
    out["tma_issue_to_full_seen_mean"] = float(lat2.mean()); out["tma_issue_to_full_seen_min"] = float(lat2.min())
    epi = t[4]
    out["epilogue_pass1_mean"] = float((epi[1:6, 1] - epi[1:6, 0]).mean()); out["epilogue_stores_mean"] = float((epi[1:6, 2] - epi[1:6, 1]).mean())
    st5 = t[5][:, 0]
    out["epi_block_stamps_tile2"] = [int(v - epi[2, 1]) for v in st5[16:24]]
    print(json.dumps(out), flush=True)
    ops.profile_enable(True); ops.profile_collect()
    for _ in range(5): h.forward(x)
    prof = ops.profile_collect(); ops.profile_enable(False)
    print("  gemm_ms", round(prof[1][0] / prof[1][1], 3))
    if False:
        for i in range(128, 140):
            print("  kb", i, "prod wait", prod[i, 0] - mma[128, 0], prod[i, 1] - mma[128, 0], "| mma wait", mma[i, 0] - mma[128, 0], mma[i, 1] - mma[128, 0], "commit", mma[i, 2] - mma[128, 0])
        for c in range(64, 70):
            print("  chunk", c, "mma tempty wait", chunk[c, 0] - mma[128, 0], chunk[c, 1] - mma[128, 0], "| fold wait", fold[c, 0] - mma[128, 0], fold[c, 1] - mma[128, 0], "done", fold[c, 2] - mma[128, 0])
