import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
from oracle import lcrec_oracle as O
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(5)
n_items, d, k = 6000, 32, 256
resid = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
resid[100:140] = resid[100] + 1e-4 * rng.standard_normal((40, d)).astype(np.float32)
resid[200:204] = resid[200]
cb = (rng.standard_normal((k, d)) * 0.1).astype(np.float32)
sizes = [2, 3, 2, 5, 8, 9, 24, 3, 2, 100, 101, 130, 2, 7, 300]
rest = rng.permutation(np.setdiff1d(np.arange(n_items), np.r_[100:140, 200:204]))
groups, pos = [np.arange(100, 140), np.arange(200, 204)], 0
for s in sizes:
    groups.append(np.sort(rest[pos:pos + s])); pos += s
mem = np.concatenate(groups).astype(np.int64)
off = np.cumsum([0] + [len(x) for x in groups]).astype(np.int64)
out = {}
for mode in (0, 1, 2):
    ops.sinkhorn_set_mode(mode)
    codes = torch.full((n_items, 4), 7, dtype=torch.int64, device=dev)
    fl, risky = ops.sinkhorn_groups(T(resid), T(cb), T(off), T(mem), torch.tensor([len(groups)], device=dev), len(groups), int(off[-1]), 0.003, 50, codes, 3, want_risky=True)
    out[mode] = codes.cpu().numpy()[:, 3]
    print("mode", mode, "risky groups", risky)
ops.sinkhorn_set_mode(2)
for gi, g in enumerate(groups):
    for m in (1, 2):
        bad = np.nonzero(out[m][g] != out[0][g])[0]
        if len(bad):
            dd = O.vq_distances(resid[g], cb); dc = O.center_distance_for_constraint(dd).astype(np.float64)
            q = O.sinkhorn_algorithm(dc, 0.003, 50)
            for i in bad:
                a, b = q[i, out[0][g][i]], q[i, out[m][g][i]]
                print(f"mode {m} group {gi} n={len(g)} row {i}: literal {out[0][g][i]} q={a!r} other {out[m][g][i]} q={b!r} rel gap {(a-b)/a:.3e} share {a*k/len(g):.6f}")
