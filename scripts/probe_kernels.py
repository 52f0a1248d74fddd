"""Bring-up probe (run on the GPU box): numerical error + timing of each kernel, printed as JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops  # noqa: E402
from oracle import lcrec_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def emit(**kw):
    print(json.dumps(kw), flush=True)


def time_fn(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_case(n, k, m, variant, chunk, relu=True, seed=0, timing=False):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, k, device=dev, generator=g)
    w = torch.randn(m, k, device=dev, generator=g) * (2.0 / (k + m)) ** 0.5
    b = torch.randn(m, device=dev, generator=g) * 0.01
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    try:
        y = ops.linear_forward(x, w, b, relu, acc_chunk=chunk, variant=variant)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        emit(kind="gemm", n=n, k=k, m=m, variant=variant, chunk=chunk, error=str(e)[:300])
        return
    t32 = torch.nn.functional.linear(x, w, b)
    if relu:
        t32 = t32.clamp_min(0)
    scale = ref.abs().mean().item()
    err = (y.double() - ref)
    err32 = (t32.double() - ref)
    out = dict(kind="gemm", n=n, k=k, m=m, variant=variant, chunk=chunk,
               rel_err=float(err.abs().mean().item() / scale), max_err=float(err.abs().max().item() / scale),
               bias=float(err.mean().item() / scale),
               torch_fp32_rel_err=float(err32.abs().mean().item() / scale),
               torch_fp32_max=float(err32.abs().max().item() / scale))
    if timing:
        ms = time_fn(lambda: ops.linear_forward(x, w, b, relu, acc_chunk=chunk, variant=variant))
        ms_t = time_fn(lambda: torch.nn.functional.linear(x, w, b))
        out.update(ms=ms, tflops_eff=2.0 * n * k * m / ms / 1e9, torch_ms=ms_t, torch_tflops=2.0 * n * k * m / ms_t / 1e9)
    emit(**out)


def main():
    emit(kind="device", name=torch.cuda.get_device_name(0), cap=torch.cuda.get_device_capability(0))
    if "--skip-gemm" not in sys.argv:
        gemm_suite()
    rest()


def gemm_suite():
    # 1. GEMM bring-up: small first (a hang/trap here is cheap), then the real shapes
    for (n, k, m) in [(128, 32, 32), (128, 64, 64), (200, 96, 64), (300, 128, 128), (257, 256, 256), (1000, 512, 512)]:
        for variant in (0, 1):
            gemm_case(n, k, m, variant, 0)
    gemm_case(300, 100, 48 + 0, 0, 0)       # K, N not multiples of the tile (N must be %4)
    for chunk in (0, 32, 64, 128, 256, 1024):
        for variant in (0, 1):
            gemm_case(2048, 4096, 2048, variant, chunk)
    for chunk in (0, 64, 256):
        for variant in (0, 1):
            gemm_case(16384, 4096, 2048, variant, chunk, timing=True)
    gemm_case(65536, 4096, 2048, 0, 0, timing=True)
    gemm_case(65536, 2048, 1024, 0, 0, timing=True)
    gemm_case(65536, 64, 32, 0, 0, relu=False, timing=True)



def rest():
    # 2. fused RQ vs oracle
    rng = np.random.default_rng(3)
    for (n, d, ks) in [(5000, 32, [256] * 4), (3000, 16, [32] * 4), (2000, 64, [128, 64, 32]), (500, 48, [100, 50])]:
        z = rng.standard_normal((n, d)).astype(np.float32)
        cbs = [(rng.standard_normal((k, d)) * 0.7 * 0.6 ** l).astype(np.float32) for l, k in enumerate(ks)]
        p = O.RqvaeParams(encoder=None, codebooks=cbs, sk_epsilons=[0.0] * len(ks))
        xq_o, loss_o, codes_o = O.rq_forward(z, p, use_sk=False)
        try:
            r = ops.rq_quantize(torch.from_numpy(z).to(dev), [torch.from_numpy(c).to(dev) for c in cbs],
                                resid_level=len(ks) - 1, want_xq=True, want_sq_err=True)
            codes = r["codes"].cpu().numpy()
            near, hard = O.classify_code_mismatches(z, p, codes)
            resids, _, _ = O.rq_trace(z, p)
            emit(kind="rq", n=n, d=d, ks=ks, mismatch_rows=int((codes != codes_o).any(1).sum()), near=near, hard=hard,
                 xq_maxdiff=float(np.abs(r["xq"].cpu().numpy() - xq_o).max()),
                 resid_maxdiff=float(np.abs(r["resid"].cpu().numpy() - resids[-1]).max()),
                 sq_err=r["sq_err"].cpu().numpy().tolist())
        except Exception as e:  # noqa: BLE001
            emit(kind="rq", n=n, d=d, error=str(e)[:300])
    zt = torch.randn(1 << 20, 32, device=dev)
    cbt = [torch.randn(256, 32, device=dev) * 0.6 ** l for l in range(4)]
    try:
        ms = time_fn(lambda: ops.rq_quantize(zt, cbt, resid_level=3))
        emit(kind="rq_time", n=1 << 20, ms=ms, items_per_s=(1 << 20) / ms * 1e3)
    except Exception as e:  # noqa: BLE001
        emit(kind="rq_time", error=str(e)[:300])

    # 3. Sinkhorn dense vs golden
    gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "sinkhorn_kat.npz"))
    for ci, (n, k, eps, iters) in enumerate(gold["meta"]):
        dc = torch.from_numpy(gold[f"dc_{ci}"]).double().to(dev)
        try:
            q, arg, flags = ops.sinkhorn_dense(dc, float(eps), int(iters), want_argmax=True)
            qr = gold[f"q_{ci}"]
            qq = q.cpu().numpy()[: qr.shape[0]]
            emit(kind="sk_dense", case=ci, n=int(n), k=int(k), argmax_mismatch=int((arg.cpu().numpy() != gold[f"arg_{ci}"]).sum()),
                 q_rel=float(np.max(np.abs(qq - qr) / np.maximum(np.abs(qr), 1e-300))), flags=int(flags.item()))
            d32 = torch.from_numpy(gold[f"d_{ci}"]).to(dev)
            dc2 = ops.center_distances(d32)
            emit(kind="centre", case=ci, maxdiff=float((dc2.cpu().numpy() - gold[f"dc_{ci}"]).__abs__().max()))
        except Exception as e:  # noqa: BLE001
            emit(kind="sk_dense", case=ci, error=str(e)[:300])
    dcb = torch.rand(1024, 256, device=dev, dtype=torch.float64) * 2 - 1
    try:
        ms = time_fn(lambda: ops.sinkhorn_dense(dcb, 0.003, 50, want_argmax=True))
        emit(kind="sk_dense_time", b=1024, k=256, ms=ms)
    except Exception as e:  # noqa: BLE001
        emit(kind="sk_dense_time", error=str(e)[:300])

    # 4. collisions + group Sinkhorn vs oracle
    for n, k, L in [(10000, 16, 3), (100000, 256, 4), (1 << 20, 256, 4)]:
        codes = rng.integers(0, k, size=(n, L)).astype(np.int64)
        dup = rng.integers(0, n, size=n // 3)
        codes[dup] = codes[rng.integers(0, n, size=n // 3)]
        ct = torch.from_numpy(codes).to(dev)
        try:
            r = ops.collisions(ct, [k] * L)
            grp = O.collision_groups(codes) if n <= 100000 else None
            ok = None
            if grp is not None:
                off = r["offsets"].cpu().numpy(); mem = r["members"].cpu().numpy()
                ours = sorted(tuple(mem[off[g]:off[g + 1]].tolist()) for g in range(r["n_groups"]))
                ok = ours == sorted(tuple(g) for g in grp)
            ms = time_fn(lambda: ops.collisions(ct, [k] * L), iters=3, warm=1)
            emit(kind="collide", n=n, n_unique=r["n_unique"], n_unique_ref=O.n_unique_codes(codes), n_groups=r["n_groups"],
                 max_mult=r["max_multiplicity"], max_mult_ref=O.max_conflicts(codes), groups_equal=ok, ms=ms)
        except Exception as e:  # noqa: BLE001
            emit(kind="collide", n=n, error=str(e)[:300])
    # groups: build groups of random sizes, compare with oracle per group
    n_items, d, k = 4000, 32, 256
    resid = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
    resid[100:140] = resid[100] + 1e-4 * rng.standard_normal((40, d)).astype(np.float32)   # near-duplicates
    resid[200:203] = resid[200]                                                             # exact duplicates
    cb = (rng.standard_normal((k, d)) * 0.1).astype(np.float32)
    sizes = [2, 3, 2, 5, 8, 9, 24, 40, 3, 2, 100, 130, 2, 7]
    rest = rng.permutation(np.setdiff1d(np.arange(n_items), np.arange(100, 140)))
    groups, pos = [], 0
    for gi, s_ in enumerate(sizes):
        if gi == 7:
            groups.append(np.arange(100, 140))       # the near-duplicate block
        else:
            groups.append(np.sort(rest[pos:pos + s_])); pos += s_
    groups[0] = np.array([200, 201])                 # exact duplicates
    mem = np.concatenate(groups).astype(np.int64)
    off = np.cumsum([0] + [len(g) for g in groups]).astype(np.int64)
    codes = np.zeros((n_items, 4), dtype=np.int64)
    ct = torch.from_numpy(codes).to(dev)
    ng = torch.tensor([len(groups)], dtype=torch.int64, device=dev)
    try:
        fl = ops.sinkhorn_groups(torch.from_numpy(resid).to(dev), torch.from_numpy(cb).to(dev), torch.from_numpy(off).to(dev),
                                 torch.from_numpy(mem).to(dev), ng, len(groups), int(off[-1]), 0.003, 50, ct, 3)
        got = ct.cpu().numpy()[:, 3]
        bad = 0; rows = 0
        for g in groups:
            idx = O.vq_assign(resid[g], cb, True, 0.003, 50)
            bad += int((got[g] != idx).sum()); rows += len(g)
        emit(kind="sk_groups", rows=rows, mismatch=bad, flags=fl)
    except Exception as e:  # noqa: BLE001
        emit(kind="sk_groups", error=str(e)[:300])


if __name__ == "__main__":
    main()
