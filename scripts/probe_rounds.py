"""Per-round breakdown of the collision loop on the bench workload (GPU box)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lcrec_b200 import ops
from lcrec_b200.models import RQVAE
from lcrec_b200 import generate_indices as G
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ws, bs, cbs, head = bench.make_model()
model = RQVAE(in_dim=4096, num_emb_list=bench.N_CODES, e_dim=32, layers=bench.DIMS[1:-1], sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
sd = model.state_dict()
lin = sorted([k for k in sd if k.startswith("encoder.mlp_layers.") and k.endswith(".weight")], key=lambda s: int(s.split(".")[2]))
for k, w, b in zip(lin, ws, bs):
    sd[k] = torch.from_numpy(w); sd[k.replace(".weight", ".bias")] = torch.from_numpy(b)
for l, cb in enumerate(cbs):
    sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
model.load_state_dict(sd); model = model.to(dev).eval()
x = bench.make_items_device(n, head, dev, 0)
ix = G.build_indexer(model, n)
ix.pass0(x)
codes0 = ix.codes_view(n).clone(); resid = ix.resid_view(n).clone()
cbt = model.rq.vq_layers[-1].embedding.weight.detach()
for mode in [int(m) for m in os.environ.get("SK_MODES", "2,1,0").split(",")]:
    ops.sinkhorn_set_mode(mode)
    codes = codes0.clone()
    tot = 0.0
    for rnd in range(21):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info = ops.collisions(codes, bench.N_CODES)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if info["n_unique"] == n or rnd == 20:
            break
        sizes = (info["offsets"][1:] - info["offsets"][:-1])
        hist = torch.bincount(sizes.clamp(max=12), minlength=13).tolist()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        fl, risky = ops.sinkhorn_groups(resid, cbt, info["offsets"], info["members"], info["counts_dev"][1:2], info["n_groups"],
                                        info["n_rows"], 0.003, 50, codes, 3, want_risky=True)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        tot += t3 - t2
        if rnd < 4 or rnd == 19:
            print(json.dumps(dict(mode=mode, round=rnd, groups=info["n_groups"], rows=info["n_rows"], risky=risky, size_hist_2_to_12plus=hist[2:],
                                  ms_collide=(t1 - t0) * 1e3, ms_sinkhorn=(t3 - t2) * 1e3)), flush=True)
    print(json.dumps(dict(mode=mode, total_sinkhorn_ms=tot * 1e3, n_unique=info["n_unique"])), flush=True)
    if mode == 0:
        lit = codes.clone()
    elif mode == 2:
        hyb = codes.clone()
if "lit" in dir() and "hyb" in dir(): print("filtered final codes == literal final codes:", bool(torch.equal(lit, hyb)))
ops.sinkhorn_set_mode(2)
