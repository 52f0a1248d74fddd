"""Probe (GPU box): classify fast-vs-literal group-Sinkhorn mismatches and time the size classes."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
from oracle import lcrec_oracle as O
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

def run(resid_items, cb, off, mem, mode):
    ops.sinkhorn_set_mode(mode)
    n_items = resid_items.shape[0]
    codes = torch.zeros((n_items, 4), dtype=torch.int64, device=dev)
    args = (T(resid_items), T(cb), T(off), T(mem), torch.tensor([len(off) - 1], device=dev), len(off) - 1, int(off[-1]), 0.003, 50, codes, 3)
    _, risky = ops.sinkhorn_groups(*args, want_risky=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.sinkhorn_groups(*args); e1.record(); torch.cuda.synchronize()
    ops.sinkhorn_set_mode(2)
    return codes.cpu().numpy()[:, 3], e0.elapsed_time(e1), risky

def make(sizes, d, k, seed, noise=0.003):
    rng = np.random.default_rng(seed)
    n_items = int(sizes.sum())
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    centres = (rng.standard_normal((len(sizes), d)) * 0.05).astype(np.float32)
    resid = np.repeat(centres, sizes, axis=0) + (rng.standard_normal((n_items, d)) * noise).astype(np.float32)
    cb = (rng.standard_normal((k, d)) * 0.05).astype(np.float32)
    mem = np.arange(n_items, dtype=np.int64)
    return resid, cb, off, mem

def classify(resid, cb, off, mem, lit, fast, tag):
    bad = np.nonzero(lit != fast)[0]
    out = []
    for r in bad[:20]:
        g = np.searchsorted(off, r, side="right") - 1
        rows = mem[off[g]:off[g + 1]]
        dd = O.vq_distances(resid[rows], cb); dc = O.center_distance_for_constraint(dd).astype(np.float64)
        q = O.sinkhorn_algorithm(dc, 0.003, 50)
        i = int(np.nonzero(rows == r)[0][0])
        a, b = q[i, lit[r]], q[i, fast[r]]
        out.append(dict(n=len(rows), lit=int(lit[r]), fast=int(fast[r]), q_lit=a, q_fast=b, ulps=abs(a - b) / np.spacing(max(a, b)),
                        oracle=int(np.argmax(q[i])), b_over_k=len(rows) / cb.shape[0]))
    print(json.dumps(dict(kind="mismatch", tag=tag, total_rows=int(len(lit)), n_bad=int(len(bad)), detail=out)), flush=True)

for (lo, hi, ng, noise) in [(2, 2, 200000, 0.003), (3, 4, 100000, 0.003), (5, 8, 60000, 0.003), (9, 24, 20000, 0.003),
                            (2, 8, 150000, 1e-4), (2, 8, 150000, 1e-6), (2, 8, 150000, 0.0)]:
    rng = np.random.default_rng(lo)
    sizes = rng.integers(lo, hi + 1, size=ng)
    resid, cb, off, mem = make(sizes, 32, 256, lo, noise)
    lit, t_lit, _ = run(resid, cb, off, mem, 0)
    fast, t_fast, _ = run(resid, cb, off, mem, 1)
    hyb, t_hyb, risky = run(resid, cb, off, mem, 2)
    print(json.dumps(dict(kind="class", lo=lo, hi=hi, groups=ng, rows=int(off[-1]), ms_literal=t_lit, ms_scaling=t_fast,
                          ms_filtered=t_hyb, risky_groups=risky, filtered_vs_literal_mismatch=int((hyb != lit).sum()),
                          scaling_vs_literal_mismatch=int((fast != lit).sum()))), flush=True)
    classify(resid, cb, off, mem, lit, hyb, f"filtered {lo}-{hi} noise {noise}")
