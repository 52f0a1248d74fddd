"""Parity tests proper (B200 only): every kernel and the assembled path against the oracle and
the committed golden vectors of the unmodified reference.  All calls go through the C ABI
(lcrec_b200.ops -> liblcrec_b200.so).  Integer outputs must be bit-exact (apart from counted
distance near-ties, top-2 gap < 1e-5 relative); floating point within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import lcrec_oracle as O
from tests.conftest import state_dict_of
from lcrec_b200.synth import seeded_weights, synth_items

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200 import ops
    from lcrec_b200 import generate_indices as G
    from lcrec_b200.models import RQVAE, MLPLayers, VectorQuantizer, sinkhorn_algorithm
    DEV = torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def small(bn=False):
    return RQVAE(in_dim=96, num_emb_list=[32] * 4, e_dim=16, layers=[64, 48], bn=bn,
                 sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)


def load_small(golden, name):
    g = golden(name)
    m = small(bool(g["bn"]))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state_dict_of(g).items()})
    p = O.params_from_state_dict(state_dict_of(g), g["sk_epsilons"].tolist(), int(g["sk_iters"]))
    return g, m.to(DEV).eval(), p


# ------------------------------------------------------------------ a2: linear / MLP
@pytest.mark.parametrize("n,k,m", [(1, 32, 32), (128, 64, 64), (200, 96, 64), (300, 100, 48), (257, 256, 256),
                                   (1000, 512, 512), (513, 4096, 2048), (640, 16, 96), (77, 200, 256), (40000, 1024, 512),
                                   (300, 256, 128), (70000, 128, 64), (129, 72, 128)])
@pytest.mark.parametrize("relu", [True, False])
def test_linear_matches_fp64(n, k, m, relu):
    g = torch.Generator(device=DEV).manual_seed(n * 7 + k)
    x = torch.randn(n, k, device=DEV, generator=g)
    w = torch.randn(m, k, device=DEV, generator=g) * (2.0 / (k + m)) ** 0.5
    b = torch.randn(m, device=DEV, generator=g) * 0.01
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    variants = [0, 1, 2, 3]           # tile variant x operand encoding (tf32 x3 / f16 x3)
    if (m % 256 == 0 or m in (128, 64)) and k >= 64:
        variants += [6, 7]            # CTA-pair kernel (cta_group::2, group-scaled operands; 256 / 128 / 64 wide tiles)
    for variant in variants:
        y = ops.linear_forward(x, w, b, relu, variant=variant)
        scale = (x.double().abs() @ w.double().abs().t()).mean().item() + 1e-30
        err = (y.double() - ref).abs().max().item() / scale
        assert err < 2e-6, (variant, err)           # tolerance: 1e-5 relative bar, measured ~3e-7
        # never worse than twice what torch's own fp32 GEMM manages
        t32 = torch.nn.functional.linear(x, w, b)
        t32 = t32.clamp_min(0) if relu else t32
        err32 = (t32.double() - ref).abs().mean().item()
        assert (y.double() - ref).abs().mean().item() <= max(3.0 * err32, 2e-7 * scale)


@pytest.mark.parametrize("n,k,m", [(1024, 4096, 2048), (424, 512, 256), (37, 96, 64), (16, 32, 32), (1000, 256, 512), (1024, 32, 4096)])
@pytest.mark.parametrize("relu", [True, False])
def test_linear_backward_matches_fp64(n, k, m, relu):
    """gx = g W, gw = g^T x, gb = sum g with g = gy * (y > 0): the training backward on the split-operand tensor-core
    GEMMs (pair kernel where the shape allows, batch sizes that are not multiples of 8 included) vs fp64."""
    g = torch.Generator(device=DEV).manual_seed(n + 3 * k + m)
    x = torch.randn(n, k, device=DEV, generator=g)
    w = torch.randn(m, k, device=DEV, generator=g) * (2.0 / (k + m)) ** 0.5
    b = torch.randn(m, device=DEV, generator=g) * 0.01
    gy = torch.randn(n, m, device=DEV, generator=g)
    y = ops.linear_forward(x, w, b, relu)
    gx, gw, gb = ops.linear_backward(x, w, y if relu else None, gy)
    gd = gy.double() * (y > 0).double() if relu else gy.double()
    for got, ref, a_abs, b_abs in ((gx, gd @ w.double(), gd.abs(), w.double().abs()),
                                   (gw, gd.t() @ x.double(), gd.abs().t(), x.double().abs())):
        scale = (a_abs @ b_abs).mean().item() + 1e-30
        assert (got.double() - ref).abs().max().item() / scale < 2e-6
    ref_b = gd.sum(0)
    assert (gb.double() - ref_b).abs().max().item() <= 1e-5 * gd.abs().sum(0).max().item() + 1e-30
    only_w = ops.linear_backward(x, w, None, gy, need_gx=False, need_gb=False)
    assert only_w[0] is None and only_w[2] is None and only_w[1].shape == (m, k)


def test_linear_zero_rows_and_errors():
    w = torch.randn(8, 8, device=DEV)
    assert ops.linear_forward(torch.zeros(0, 8, device=DEV), w, None, True).shape == (0, 8)
    with pytest.raises(RuntimeError):
        ops.linear_forward(torch.zeros(2, 8), w, None, True)            # CPU tensor: loud failure


def test_fullshape_encoder_matches_reference(golden):
    g = golden("fullshape")
    dims = g["dims"].tolist()
    ws, bs, cbs = seeded_weights(dims, [256] * 4, 32, seed=int(g["seed_w"]))
    x = synth_items(int(g["n"]), dims[0], n_parents=int(g["n"]) // 8, seed=int(g["seed_x"]))
    h = ops.MlpHandle([T(w) for w in ws], [T(b) for b in bs])
    z = h.forward(T(x)).cpu().numpy()
    scale = np.abs(g["latents"]).max()
    assert np.abs(z - g["latents"]).max() / scale < 1e-5
    p = O.RqvaeParams(encoder=O.MlpParams(ws, bs), codebooks=cbs, sk_epsilons=[0, 0, 0, 0.003])
    codes = ops.rq_quantize(T(z), [T(c) for c in cbs])["codes"].cpu().numpy()
    near, hard = O.classify_code_mismatches(g["latents"], p, codes)
    assert hard == 0 and near <= 2
    assert (codes != g["codes"]).any(axis=1).sum() <= 2
    # activations of hidden layers are returned on request (training path)
    y, acts = h.forward(T(x[:130]), want_acts=True)
    np.testing.assert_allclose(acts[0][:8, :64].cpu().numpy(), g["h1_head"], rtol=2e-5, atol=2e-6)
    assert torch.equal(acts[-1], y)


# ------------------------------------------------------------------ a4/a9: fused RQ
@pytest.mark.parametrize("n,d,ks", [(0, 32, [256] * 4), (1, 32, [256] * 4), (5000, 32, [256] * 4), (3000, 16, [32] * 4),
                                    (2000, 64, [128, 64, 32]), (500, 48, [100, 50]), (300, 256, [512, 512])])
def test_rq_fused_matches_oracle(n, d, ks):
    rng = np.random.default_rng(n + d)
    z = rng.standard_normal((n, d)).astype(np.float32)
    if n > 10:
        z[7] = z[3]                                   # duplicate rows must get identical codes
    cbs = [(rng.standard_normal((k, d)) * 0.7 * 0.6 ** l).astype(np.float32) for l, k in enumerate(ks)]
    cbs[0][5] = cbs[0][2]                             # exact tie between two codes: lowest index wins
    p = O.RqvaeParams(encoder=None, codebooks=cbs, sk_epsilons=[0.0] * len(ks))
    r = ops.rq_quantize(T(z), [T(c) for c in cbs], resid_level=len(ks) - 1, want_xq=True, want_sq_err=True)
    if n == 0:
        assert r["codes"].shape == (0, len(ks))
        return
    xq_o, loss_o, codes_o = O.rq_forward(z, p, use_sk=False)
    codes = r["codes"].cpu().numpy()
    near, hard = O.classify_code_mismatches(z, p, codes)
    assert hard == 0
    same = (codes == codes_o).all(axis=1)
    assert (~same).sum() == near
    assert not (codes[:, 0] == 5).any()
    np.testing.assert_allclose(r["xq"].cpu().numpy()[same], xq_o[same], rtol=1e-5, atol=1e-6)
    resids, _, _ = O.rq_trace(z, p)
    np.testing.assert_allclose(r["resid"].cpu().numpy()[same], resids[-1][same], rtol=1e-5, atol=1e-6)
    mse = r["sq_err"].cpu().numpy() / (n * d)
    loss = np.mean([(m + p.beta * m) for m in mse.astype(np.float32)])
    np.testing.assert_allclose(loss, loss_o, rtol=1e-5)


@pytest.mark.parametrize("n,d,ks,force", [(3000, 256, [8192, 8192], False), (5000, 128, [512, 256, 256], False),
                                          (6000, 32, [256] * 4, True), (2500, 64, [4096, 256], True), (1500, 40, [256, 512], True)])
def test_rq_tensor_core_path_matches_oracle(n, d, ks, force):
    """Large-codebook variant (BASELINE configs[4]: 8192 codes x 256 dims): distance GEMM + argmin on the tensor cores
    (CTA-pair tcgen05 kernel, fp32-accurate split operands).  Codes equal the oracle's apart from counted near-ties;
    forced onto the small shapes it equals the SIMT kernel the same way."""
    rng = np.random.default_rng(n + d)
    z = rng.standard_normal((n, d)).astype(np.float32)
    z[7] = z[3]
    cbs = [(rng.standard_normal((k, d)) * 0.7 * 0.6 ** l).astype(np.float32) for l, k in enumerate(ks)]
    cbs[0][5] = cbs[0][2]                             # exact tie between two codes: lowest index wins
    p = O.RqvaeParams(encoder=None, codebooks=cbs, sk_epsilons=[0.0] * len(ks))
    try:
        ops.rq_set_tc_mode(2 if force else 1)
        r = ops.rq_quantize(T(z), [T(c) for c in cbs], resid_level=len(ks) - 1, want_xq=True, want_sq_err=True)
    finally:
        ops.rq_set_tc_mode(1)
    codes = r["codes"].cpu().numpy()
    xq_o, loss_o, codes_o = O.rq_forward(z, p, use_sk=False)
    near, hard = O.classify_code_mismatches(z, p, codes)
    assert hard == 0 and near <= max(3, n // 500)
    same = (codes == codes_o).all(axis=1)
    assert (~same).sum() <= near
    assert (codes[7] == codes[3]).all() and not (codes[:, 0] == 5).any()
    np.testing.assert_allclose(r["xq"].cpu().numpy()[same], xq_o[same], rtol=1e-5, atol=1e-6)
    resids, _, _ = O.rq_trace(z, p)
    np.testing.assert_allclose(r["resid"].cpu().numpy()[same], resids[-1][same], rtol=1e-5, atol=1e-6)
    if same.all():
        mse = r["sq_err"].cpu().numpy() / (n * d)
        loss = np.mean([(m + p.beta * m) for m in mse.astype(np.float32)])
        np.testing.assert_allclose(loss, loss_o, rtol=1e-5)
    ops.rq_set_tc_mode(0)
    try:
        r0 = ops.rq_quantize(T(z), [T(c) for c in cbs])
    finally:
        ops.rq_set_tc_mode(1)
    diff = (r0["codes"].cpu().numpy() != codes).any(axis=1)
    assert diff.sum() <= near + max(3, n // 500)      # SIMT path vs tensor-core path: near-ties only


def test_vq_distances_match_oracle():
    rng = np.random.default_rng(0)
    r = rng.standard_normal((300, 32)).astype(np.float32)
    cb = rng.standard_normal((256, 32)).astype(np.float32)
    d = ops.vq_distances(T(r), T(cb)).cpu().numpy()
    np.testing.assert_allclose(d, O.vq_distances(r, cb), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ a5/a6/a7: Sinkhorn
def test_sinkhorn_dense_matches_reference_golden(golden):
    g = golden("sinkhorn_kat")
    for ci, (n, k, eps, iters) in enumerate(g["meta"]):
        dc = T(g[f"dc_{ci}"]).double()
        q, arg, flags = ops.sinkhorn_dense(dc, float(eps), int(iters), want_argmax=True)
        assert (arg.cpu().numpy() == g[f"arg_{ci}"]).all(), f"case {ci}"      # exact, incl. exact-tie rows
        qr = g[f"q_{ci}"]
        np.testing.assert_allclose(q.cpu().numpy()[: qr.shape[0]], qr, rtol=1e-9, atol=1e-300)
        assert int(flags.item()) == 0
        # centring (fp32) is bit-exact
        assert np.array_equal(ops.center_distances(T(g[f"d_{ci}"])).cpu().numpy(), g[f"dc_{ci}"].astype(np.float64))
        # module-level drop-in
        q2 = sinkhorn_algorithm(dc, float(eps), int(iters))
        assert torch.equal(q2, q)
        assert np.array_equal(VectorQuantizer.center_distance_for_constraint(T(g[f"d_{ci}"])).cpu().numpy(), g[f"dc_{ci}"])


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_sharded_sinkhorn_equals_single_gpu():
    """DP form of the training-time Sinkhorn: rows split over 2 ranks, column marginals all-reduced inside the kernel
    through peer memory; argmax identical and Q within 1e-9 of the single-GPU kernel on the concatenated batch."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", os.path.join(root, "scripts", "check_dist_sinkhorn.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"] and all(c["argmax_mismatches"] == 0 for c in out["cases"])


def test_sinkhorn_nan_flag_and_amplitude_assert():
    d = torch.zeros(4, 8, device=DEV, dtype=torch.float64)
    d[0, 0] = float("nan")
    q, arg, flags = ops.sinkhorn_dense(d, 0.003, 5, want_argmax=True)
    assert int(flags.item()) & 1
    with pytest.raises(AssertionError):
        ops.center_distances(torch.full((2, 4), float("nan"), device=DEV))


def test_sinkhorn_groups_match_oracle():
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
    codes = torch.full((n_items, 4), 7, dtype=torch.int64, device=DEV)
    fl = ops.sinkhorn_groups(T(resid), T(cb), T(off), T(mem), torch.tensor([len(groups)], device=DEV),
                             len(groups), int(off[-1]), 0.003, 50, codes, 3)
    assert fl == 0
    got = codes.cpu().numpy()
    assert (got[:, :3] == 7).all()
    untouched = np.setdiff1d(np.arange(n_items), mem)
    assert (got[untouched, 3] == 7).all()
    bad = rows = 0
    for g in groups:
        idx = O.vq_assign(resid[g], cb, True, 0.003, 50)
        bad += int((got[g, 3] != idx).sum()); rows += len(g)
    assert bad == 0, (bad, rows)
    # the same through the other two arithmetic modes
    for mode, tol in ((0, 0), (1, 3)):
        try:
            ops.sinkhorn_set_mode(mode)
            c2 = torch.full((n_items, 4), 7, dtype=torch.int64, device=DEV)
            ops.sinkhorn_groups(T(resid), T(cb), T(off), T(mem), torch.tensor([len(groups)], device=DEV),
                                len(groups), int(off[-1]), 0.003, 50, c2, 3)
        finally:
            ops.sinkhorn_set_mode(2)
        assert int((c2.cpu().numpy() != got).sum()) <= tol, mode


@pytest.mark.parametrize("k,d", [(256, 32), (32, 16), (100, 32), (512, 32)])
def test_sinkhorn_group_modes_agree(k, d):
    """The three arithmetic modes on 30k groups in the exact-tie regime (near-duplicate + duplicate rows, sizes
    2..12 plus a few large): the default filtered mode (2) must give exactly the codes of the all-literal
    mode (0); the pure scaling form (1) may differ on ulp-level ties only (counted, < 1e-4 of rows)."""
    rng = np.random.default_rng(11)
    sizes = np.concatenate([rng.integers(2, 13, size=30000), rng.integers(13, 34, size=400), [40, 99, 100, 150]])
    n_items = int(sizes.sum())
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    centres = (rng.standard_normal((len(sizes), d)) * 0.05).astype(np.float32)
    resid = np.repeat(centres, sizes, axis=0) + (rng.standard_normal((n_items, d)) * 0.003).astype(np.float32)
    dup = rng.integers(0, n_items - 1, size=2000)
    same_group = np.searchsorted(off, dup, side="right") == np.searchsorted(off, dup + 1, side="right")
    resid[dup[same_group] + 1] = resid[dup[same_group]]                 # exact duplicates inside groups
    cb = (rng.standard_normal((k, d)) * 0.05).astype(np.float32)
    mem = rng.permutation(n_items).astype(np.int64)
    for g in range(len(sizes)):
        mem[off[g]:off[g + 1]].sort()
    resid_items = np.empty_like(resid)
    resid_items[mem] = resid
    out = {}
    try:
        for mode in (0, 1, 2):
            ops.sinkhorn_set_mode(mode)
            codes = torch.zeros((n_items, 4), dtype=torch.int64, device=DEV)
            fl = ops.sinkhorn_groups(T(resid_items), T(cb), T(off), T(mem), torch.tensor([len(sizes)], device=DEV),
                                     len(sizes), n_items, 0.003, 50, codes, 3)
            assert fl == 0
            out[mode] = codes.cpu().numpy()[:, 3]
        # groups of 9..32 rows: the column kernels (default) against the shared-memory CTA kernel, filtered mode
        ops.sinkhorn_set_mode(2)
        ops.sinkhorn_set_col(False)
        codes = torch.zeros((n_items, 4), dtype=torch.int64, device=DEV)
        ops.sinkhorn_groups(T(resid_items), T(cb), T(off), T(mem), torch.tensor([len(sizes)], device=DEV),
                            len(sizes), n_items, 0.003, 50, codes, 3)
        out["cta"] = codes.cpu().numpy()[:, 3]
        # a late collision round: a call with few groups runs EVERY group on the column kernels (one thread per code, RM = 4 / 8 /
        # 16 / 32) - against the warp kernels (col = 1) and the literal mode on the same 800 groups
        sub = np.r_[0:700, len(sizes) - 100:len(sizes)]                      # sizes 2..12, 13..33 and the large ones
        sub_off = np.concatenate([[0], np.cumsum(sizes[sub])]).astype(np.int64)
        sub_mem = np.concatenate([mem[off[g]:off[g + 1]] for g in sub]).astype(np.int64)
        for col in (2, 1):
            ops.sinkhorn_set_col(col)
            codes = torch.zeros((n_items, 4), dtype=torch.int64, device=DEV)
            fl = ops.sinkhorn_groups(T(resid_items), T(cb), T(sub_off), T(sub_mem), torch.tensor([len(sub)], device=DEV),
                                     len(sub), len(sub_mem), 0.003, 50, codes, 3)
            assert fl == 0
            out[f"late{col}"] = codes.cpu().numpy()[sub_mem, 3]
        out["late_ref"] = out[0][sub_mem]
    finally:
        ops.sinkhorn_set_mode(2)
        ops.sinkhorn_set_col(True)
    assert (out["late2"] != out["late_ref"]).sum() == 0 and (out["late1"] != out["late_ref"]).sum() == 0, \
        (int((out["late2"] != out["late_ref"]).sum()), int((out["late1"] != out["late_ref"]).sum()))
    assert (out[2] != out[0]).sum() == 0, int((out[2] != out[0]).sum())
    assert (out["cta"] != out[0]).sum() == 0, int((out["cta"] != out[0]).sum())
    assert (out[1] != out[0]).mean() < 1e-4
    # spot-check against the numpy oracle on the first 300 groups, distances evaluated in the kernels' summation order (fma
    # chains): every row must agree; a differing row is COUNTED only if the oracle's own plan holds the two columns within
    # 4 ulp of each other, anything else is hard.
    counted = hard = rows_n = 0
    for g in range(300):
        rows = mem[off[g]:off[g + 1]]
        idx, _, q = O.vq_assign(resid_items[rows], cb, True, 0.003, 50, want_q=True, order="chain")
        for i in np.nonzero(out[0][rows] != idx)[0]:
            a_, b_ = q[i, idx[i]], q[i, out[0][rows][i]]
            if abs(a_ - b_) <= 4 * np.spacing(abs(a_)):
                counted += 1
            else:
                hard += 1
        rows_n += len(rows)
    assert hard == 0 and counted <= 2, (counted, hard, rows_n)


@pytest.mark.parametrize("k,d", [(2000, 144), (8192, 256)])
def test_sinkhorn_groups_large_codebook_match_oracle(k, d):
    """Large codebooks (BASELINE configs[4]: 8192 x 256): the CTA kernels stream the codebook through a shared-memory
    tile (ragged K and e_dim included; 8192 x 256 runs on the cluster path); argmax vs the numpy oracle in the kernels'
    summation order: exact, apart from rows the oracle's own plan holds within 4 ulp (counted)."""
    rng = np.random.default_rng(k + d)
    sizes = np.concatenate([rng.integers(2, 12, size=60), [20, 3, 2]])
    n_items = int(sizes.sum())
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    centres = (rng.standard_normal((len(sizes), d)) * 0.05).astype(np.float32)
    resid = np.repeat(centres, sizes, axis=0) + (rng.standard_normal((n_items, d)) * 0.01).astype(np.float32)
    cb = (rng.standard_normal((k, d)) * 0.05).astype(np.float32)
    mem = np.arange(n_items, dtype=np.int64)
    codes = torch.zeros((n_items, 4), dtype=torch.int64, device=DEV)
    fl = ops.sinkhorn_groups(T(resid), T(cb), T(off), T(mem), torch.tensor([len(sizes)], device=DEV), len(sizes), n_items,
                             0.003, 50, codes, 3)
    assert fl == 0
    got = codes.cpu().numpy()[:, 3]
    counted = hard = 0
    for g in range(len(sizes)):
        rows = mem[off[g]:off[g + 1]]
        idx, _, q = O.vq_assign(resid[rows], cb, True, 0.003, 50, want_q=True, order="chain")     # kernel summation order
        for i in np.nonzero(got[rows] != idx)[0]:
            a_, b_ = q[i, idx[i]], q[i, got[rows][i]]
            if abs(a_ - b_) <= 4 * np.spacing(abs(a_)):
                counted += 1
            else:
                hard += 1
    assert hard == 0 and counted <= 2, (counted, hard, n_items)


@pytest.mark.parametrize("k,d", [(8192, 256), (2048, 64), (4096, 128)])
@pytest.mark.parametrize("mode", [2, 1])
def test_sinkhorn_wide_path_is_exact(k, d, mode):
    """Large codebooks on the cluster path (csrc/sinkhorn_wide.cuh: batched fma-chain distances + one cluster of 1 / 2 / 4 / 8
    CTAs per group): the picks equal the oracle evaluated in the kernels' summation order on EVERY row (groups of 2 ... 24 rows
    on the clusters, 25 and 40 rows on the CTA kernel), and equal the CTA-kernel-only path."""
    rng = np.random.default_rng(k + d + mode)
    sizes = np.array([2, 3, 2, 4, 6, 7, 12, 13, 24, 25, 40, 2, 3, 5, 9, 2, 2, 3, 16, 8, 3, 2])
    n_items = int(sizes.sum())
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    centres = (rng.standard_normal((len(sizes), d)) * 0.05).astype(np.float32)
    resid = np.repeat(centres, sizes, axis=0) + (rng.standard_normal((n_items, d)) * 0.01).astype(np.float32)
    resid[off[3] + 1] = resid[off[3]]                    # exact duplicates inside a group: the exact-tie regime
    cb = (rng.standard_normal((k, d)) * 0.05).astype(np.float32)
    mem = rng.permutation(n_items + 7)[:n_items].astype(np.int64)          # members are not the identity
    rows_all = np.zeros((n_items + 7, d), np.float32)
    rows_all[mem] = resid
    out = {}
    try:
        ops.sinkhorn_set_mode(mode)
        for wide in (True, False, 2, 3):               # cluster path, CTA kernel only, cluster path literal form / shared-memory kernels only
            ops.sinkhorn_set_wide(wide)
            codes = torch.zeros((n_items + 7, 4), dtype=torch.int64, device=DEV)
            fl = ops.sinkhorn_groups(T(rows_all), T(cb), T(off), T(mem), torch.tensor([len(sizes)], device=DEV), len(sizes), n_items,
                                     0.003, 50, codes, 3)
            assert fl == 0
            out[wide] = codes.cpu().numpy()[:, 3]
    finally:
        ops.sinkhorn_set_mode(2)
        ops.sinkhorn_set_wide(True)
    # every variant against the oracle in the kernels' distance order; a differing row is COUNTED only if the oracle's own plan
    # holds the two columns within 4 ulp of each other (the summation order inside the fp64 row / column sums is the
    # library's in the reference, so a last-ulp near-tie of Q may resolve either way), anything else is hard
    report = {}
    for name, got in out.items():
        counted = hard = 0
        for g in range(len(sizes)):
            rows = mem[off[g]:off[g + 1]]
            want, _, q = O.vq_assign(rows_all[rows], cb, True, 0.003, 50, want_q=True, order="chain")
            for i in np.nonzero(got[rows] != want)[0]:
                a_, b_ = q[i, want[i]], q[i, got[rows][i]]
                if abs(a_ - b_) <= 4 * np.spacing(abs(a_)):
                    counted += 1
                else:
                    hard += 1
        report[name] = (counted, hard)
    print(f"\n[wide k={k} d={d} mode={mode}] (counted, hard) per variant {report}; cluster vs CTA-only rows differing "
          f"{int((out[True] != out[False]).sum())}, cluster vs cluster-literal {int((out[True] != out[2]).sum())}")
    assert all(h == 0 for _, h in report.values()), report
    assert all(c <= 2 for c, _ in report.values()), report


# ------------------------------------------------------------------ a12/a14: collisions
@pytest.mark.parametrize("n,k,L", [(0, 16, 3), (1, 16, 3), (5, 4, 2), (10000, 16, 3), (100000, 256, 4), (70000, 8192, 4),
                                   (4097, 65536, 4)])
def test_collisions_match_oracle(n, k, L):
    rng = np.random.default_rng(n + k)
    codes = rng.integers(0, k, size=(n, L)).astype(np.int64)
    if n > 10:
        src = rng.integers(0, n, size=n // 3)
        codes[rng.integers(0, n, size=n // 3)] = codes[src]
        codes[-1] = k - 1                              # maximum code value in every level
    r = ops.collisions(T(codes), [k] * L)
    if n == 0:
        assert r["n_groups"] == 0
        return
    assert r["n_unique"] == O.n_unique_codes(codes)
    assert r["max_multiplicity"] == O.max_conflicts(codes)
    grp = O.collision_groups(codes)
    assert r["n_groups"] == len(grp)
    off, mem = r["offsets"].cpu().numpy(), r["members"].cpu().numpy()
    ours = sorted(tuple(mem[off[g]:off[g + 1]].tolist()) for g in range(r["n_groups"]))
    assert ours == sorted(tuple(g) for g in grp)       # members ascending inside every group
    keys, items = ops.sort_codes(T(codes), [k] * L)
    kk = keys.cpu().numpy().astype(np.uint64)
    assert (np.diff(kk.astype(np.float64)) >= 0).all() and sorted(items.cpu().tolist()) == list(range(n))


@pytest.mark.parametrize("n,k,L,seg", [(5, 4, 2, 2), (20000, 16, 3, 6), (200000, 256, 4, 5), (60000, 64, 4, 40), (9000, 8, 3, 700),
                                       (5000, 4, 2, 1500)])
def test_collisions_in_prefix_segments_match_global_sort(n, k, L, seg):
    """The per-round collision search inside prefix segments == the global sort (groups as sets of member tuples),
    for tiny segments (warp sort), segments of 33..1024 items (CTA sort) and the > 1024 fallback flag."""
    rng = np.random.default_rng(n + seg)
    n_par = max(n // seg, 1)
    prefix = rng.integers(0, k, size=(n_par, L - 1))
    codes = np.concatenate([prefix[rng.integers(0, n_par, size=n)], rng.integers(0, k, size=(n, 1))], axis=1).astype(np.int64)
    segs = ops.prefix_segments(T(codes), [k] * L)
    pre = O.collision_groups(codes[:, :-1])
    assert segs["n_segments"] == len(pre) and segs["n_rows"] == sum(len(g) for g in pre)
    assert segs["max_size"] == max([len(g) for g in pre] + [1 if n else 0])
    for rnd in range(3):                                # the last level changes between rounds, the segments stay
        codes[:, -1] = rng.integers(0, max(k // (rnd + 1), 2), size=n)
        r = ops.collisions_in_segments(T(codes), segs)
        if segs["max_size"] > 1024:
            assert r["fallback"]
            continue
        assert not r["fallback"]
        ref = ops.collisions(T(codes), [k] * L)
        for key in ("n_unique", "n_groups", "n_rows", "max_multiplicity"):
            assert r[key] == ref[key], key
        def as_set(d):
            off, mem = d["offsets"].cpu().numpy(), d["members"].cpu().numpy()
            return sorted(tuple(mem[off[g]:off[g + 1]].tolist()) for g in range(d["n_groups"]))
        assert as_set(r) == as_set(ref)
        assert int(r["offsets"][-1]) == r["n_rows"]


def test_sort_is_a_permutation_at_full_size():
    """Size-independent properties at 10M items: sortedness, permutation, unique count vs torch.unique."""
    n, k, L = 10_000_000, 256, 4
    g = torch.Generator(device=DEV).manual_seed(1)
    codes = torch.randint(0, k, (n, L), device=DEV, generator=g)
    codes[: n // 4] = codes[n // 4: n // 2]
    keys, items = ops.sort_codes(codes, [k] * L)
    assert bool((keys[1:] >= keys[:-1]).all())
    assert int(torch.bincount(items.long(), minlength=n).max()) == 1
    packed = (codes[:, 0] << 24) | (codes[:, 1] << 16) | (codes[:, 2] << 8) | codes[:, 3]
    assert torch.equal(packed[items.long()], keys)
    same = keys[1:] == keys[:-1]
    assert bool((items[1:][same] > items[:-1][same]).all())        # stable: ascending item ids in a run
    r = ops.collisions(codes, [k] * L)
    assert r["n_unique"] == int(torch.unique(packed).numel())


# ------------------------------------------------------------------ a10/a15: models and generation
@pytest.mark.parametrize("name", ["small_model", "bn_model"])
def test_model_matches_reference(golden, name):
    g, m, p = load_small(golden, name)
    x = T(g["x"])
    with torch.no_grad():
        z = m.encoder(x)
        np.testing.assert_allclose(z.cpu().numpy(), g["latents"], rtol=2e-5, atol=2e-6)
        codes = m.get_indices(x, use_sk=False).cpu().numpy()
        near, hard = O.classify_code_mismatches(g["latents"], p, codes)
        assert hard == 0 and near <= 2
        assert (codes != g["codes_argmin_full"]).any(axis=1).sum() <= 2
        xq, loss, idx = m.rq(T(g["latents"]), use_sk=False)               # teacher-forced on reference latents
        np.testing.assert_allclose(xq.cpu().numpy(), g["rq_xq"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(loss.item(), g["rq_loss"], rtol=1e-5)
        # use_sk=True (the training forward, vq.py:76-83: one dense Sinkhorn problem on the last level).  Teacher-forced on
        # the REFERENCE latents of those rows: the residual quantiser must equal the oracle evaluated in the kernels' fp32
        # summation order exactly (codes) / within 1e-5 (x_q, loss); against the reference's own indices (torch / MKL
        # summation order, encoder re-run at batch 512) the differing rows are COUNTED: at most 2 of 512, each one a row whose
        # pick the oracle reproduces on the same latents.
        lat = g["latents"][:512]
        xq_s, loss_s, idx_s = m.rq(T(lat), use_sk=True)
        oxq, oloss, oidx = O.rq_forward(lat, p, use_sk=True, order="chain")
        assert (idx_s.cpu().numpy() == oidx).all()
        np.testing.assert_allclose(xq_s.cpu().numpy(), oxq, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(loss_s.item(), float(oloss), rtol=1e-5)
        counted_tf = int((oidx != g["fwd_idx"]).any(axis=1).sum())
        assert counted_tf <= 2, counted_tf
        if counted_tf == 0:
            np.testing.assert_allclose(loss_s.item(), g["fwd_rq_loss"], rtol=1e-5)
        out, rq_loss, fidx = m(x[:512], use_sk=True)                       # the whole forward from the embeddings
        tot, rec = m.compute_loss(out, rq_loss, xs=x[:512])
    counted = int((fidx.cpu().numpy() != g["fwd_idx"]).any(axis=1).sum())
    print(f"\n[{name}] use_sk forward: rows differing from the reference: teacher-forced on its latents {counted_tf}/512, "
          f"from the embeddings {counted}/512")
    assert counted <= 2, counted
    tol = 1e-5 if counted == 0 else 1e-4          # a differing row moves a mean over 512 rows by up to ~1/512 of its term
    np.testing.assert_allclose(rec.item(), g["fwd_recon"], rtol=tol)
    np.testing.assert_allclose(tot.item(), g["fwd_total"], rtol=tol)
    # VectorQuantizer.get_code alias
    vq = m.rq.vq_layers[0]
    assert torch.equal(vq.get_code(z, use_sk=False), T(codes[:, 0]))


@pytest.mark.parametrize("name", ["small_model", "bn_model"])
def test_generate_indices_rounds_teacher_forced(golden, name):
    """Round t+1 of the collision loop from the reference's round-t table (generate_indices.py:107-128)."""
    g, m, p = load_small(golden, name)
    x = T(g["x"])
    n = x.shape[0]
    ix = G.build_indexer(m, n)
    ix.pass0(x)
    pass0 = ix.codes_view(n).cpu().numpy().copy()
    assert (pass0 != g["codes_pass0"]).any(axis=1).sum() <= 2
    tables = [g["codes_pass0"]] + list(g["rounds"])
    resid = ix.resid_view(n).cpu().numpy().copy()
    cb_last = p.codebooks[-1]
    eps = O.generation_epsilons(p)[-1]
    counted = hard = rows = 0
    for t in range(len(tables) - 1):
        ix.codes_view(n).copy_(T(tables[t]))
        c = ix.round(n)
        grp = O.collision_groups(tables[t])
        assert c["n_groups"] == len(grp) and c["n_rows"] == sum(len(v) for v in grp)
        assert c["n_unique"] == O.n_unique_codes(tables[t])
        got = ix.codes_view(n).cpu().numpy()
        rows += c["n_rows"]
        bad = np.nonzero(got[:, -1] != tables[t + 1][:, -1])[0]
        owner = {i: gi for gi, gg in enumerate(grp) for i in gg}
        for gi in sorted({owner[int(i)] for i in bad}):
            gg = grp[gi]
            # counted = the kernels are exact on their own rows (oracle in the kernels' summation order makes the same pick);
            # the difference to the reference then comes from the fp32 rounding of the rows (encoder GEMMs, 1e-5 bar)
            exact = bool((O.vq_assign(resid[gg], cb_last, True, eps, p.sk_iters, order="chain") == got[gg, -1]).all())
            k = int(np.isin(bad, gg).sum())
            counted += k if exact else 0
            hard += 0 if exact else k
    print(f"\n[{name}] teacher-forced rounds: {rows} Sinkhorn rows, counted {counted}, hard {hard}")
    assert hard == 0
    assert counted <= max(2, rows // 500), (counted, rows)


@pytest.mark.parametrize("name", ["small_model", "bn_model"])
def test_generate_indices_end_to_end(golden, name, tmp_path):
    g, m, p = load_small(golden, name)
    ref_final = g["script_codes_final"]
    n = ref_final.shape[0]
    codes_dev, stats_dev = G.generate_codes(m, T(g["x"]))                  # device-resident input
    codes_host, stats_host = G.generate_codes(m, g["x"], chunk_rows=512)   # host input, streamed in 4 chunks
    assert torch.equal(codes_dev, codes_host)
    assert stats_dev["n_unique"] == O.n_unique_codes(codes_dev.numpy())
    assert stats_dev["max_multiplicity"] == O.max_conflicts(codes_dev.numpy())
    assert abs(stats_dev["collision_rate"] - O.collision_rate(ref_final)) < 5e-3
    # LOOP exact on its own inputs (oracle from the GPU's PASS-0 table + residual rows, kernel summation order) ...
    ix = G.build_indexer(m, n)
    ix.pass0(T(g["x"]))
    c0, r3 = ix.codes_view(n).cpu().numpy().copy(), ix.resid_view(n).cpu().numpy().copy()
    want, tr = O.resolve_collisions(c0, r3, p.codebooks[-1], O.generation_epsilons(p)[-1], p.sk_iters, order="chain")
    assert (codes_dev.numpy() == want).all()
    # ... and against the reference's own run a ledger: primary rows (same group, same incoming table, other pick) + cascades
    led = O.loop_ledger([g["codes_pass0"]] + list(g["rounds"]), [c0] + tr.rounds)
    print(f"\n[{name}] end to end vs the reference script: {led.rows_differing_final}/{n} rows differ = {len(led.primary)} primary "
          f"(+ PASS-0 near-ties {led.per_round[0]}) and {led.cascade} cascade rows; per round {led.per_round}")
    assert led.rows_differing_final <= max(4, n // 100) and len(led.primary) <= 6
    assert (codes_dev.numpy()[:, :3] != ref_final[:, :3]).any(axis=1).sum() <= 2
    G.write_index_json(codes_dev, str(tmp_path / "a.json"))
    assert (tmp_path / "a.json").read_text() == O.index_json(codes_dev.numpy())


def test_segment_rounds_equal_global_sort_rounds():
    """Whole generation with and without the prefix-segment shortcut gives identical codes and statistics."""
    dims = [4096, 2048, 1024, 512, 256, 128, 64, 32]
    ws, bs, cbs = seeded_weights(dims, [256] * 4, 32, seed=3, cb_scale=0.02)
    x = T(synth_items(20000, 4096, n_parents=2500, seed=11))
    m = RQVAE(in_dim=4096, num_emb_list=[256] * 4, e_dim=32, layers=dims[1:-1], sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
    sd = m.state_dict()
    lin = sorted([k for k in sd if k.startswith("encoder.mlp_layers.") and k.endswith(".weight")], key=lambda s: int(s.split(".")[2]))
    for k, w, b in zip(lin, ws, bs):
        sd[k] = torch.from_numpy(w); sd[k.replace(".weight", ".bias")] = torch.from_numpy(b)
    for l, cb in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    try:
        ops.indexer_set_segments(False)
        c0, s0 = G.generate_codes(m, x)
        ops.indexer_set_segments(True)
        c1, s1 = G.generate_codes(m, x)
    finally:
        ops.indexer_set_segments(True)
    assert s0["rounds"] >= 2 and s0 == s1
    assert torch.equal(c0, c1)
    # the rounds alone on caller-owned tensors (the multi-GPU bucket path) reproduce the same table
    ix = G.build_indexer(m, x.shape[0])
    ix.pass0(x)
    codes, resid = ix.codes_view(x.shape[0]).clone(), ix.resid_view(x.shape[0]).clone()
    s2 = ix.resolve_device(codes, resid, 20)
    assert torch.equal(codes.cpu(), c1.cpu()) and s2["n_unique"] == s1["n_unique"]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_generation_equals_single_gpu_bitwise():
    """2 ranks (NCCL all-to-all bucket exchange + native rounds per owner) == the single-GPU table, bit for bit."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29578", os.path.join(root, "scripts", "check_sharded_equals_single.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"] and out["identical"]


def test_full_size_generation_properties():
    """BASELINE configs[2] size (1 M items x 4096): size-independent properties of the whole path - determinism, the
    statistics describe the table, items whose PASS-0 code was unique are untouched, a second resolve of the final
    table only touches what still collides, and the prefix (first L-1 levels) never changes."""
    import bench
    ws, bs, cbs, head = bench.make_model()
    m = RQVAE(in_dim=4096, num_emb_list=bench.N_CODES, e_dim=32, layers=bench.DIMS[1:-1], sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
    sd = m.state_dict()
    lin = sorted([k for k in sd if k.startswith("encoder.mlp_layers.") and k.endswith(".weight")], key=lambda s: int(s.split(".")[2]))
    for k, w, b in zip(lin, ws, bs):
        sd[k] = torch.from_numpy(w); sd[k.replace(".weight", ".bias")] = torch.from_numpy(b)
    for l, cb in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    n = 1_000_000
    x = bench.make_items_device(n, head, DEV, 0)
    ix = G.build_indexer(m, n)
    c1, s1 = ix.run_device(x, 20)
    c2, s2 = ix.run_device(x, 20)
    assert torch.equal(c1, c2) and s1 == s2                                   # deterministic
    ix.pass0(x)
    pass0 = ix.codes_view(n).clone()
    assert torch.equal(pass0[:, :3], c1[:, :3])                               # the rounds rewrite the last level only
    key = lambda c: (c[:, 0] << 24) | (c[:, 1] << 16) | (c[:, 2] << 8) | c[:, 3]
    k0, k1 = key(pass0), key(c1)
    assert int(torch.unique(k1).numel()) == s1["n_unique"]                    # the statistics describe the table
    u, inv, cnt = torch.unique(k0, return_inverse=True, return_counts=True)
    lonely = cnt[inv] == 1
    assert s1["rows_round1"] == int((~lonely).sum()) and s1["groups_round1"] == int((cnt > 1).sum())
    prefix_cnt = torch.unique(k0 >> 8, return_inverse=True, return_counts=True)
    alone_in_prefix = prefix_cnt[2][prefix_cnt[1]] == 1
    assert torch.equal(pass0[alone_in_prefix], c1[alone_in_prefix])           # nobody to collide with: never touched
    assert (c1[:, 3] >= 0).all() and (c1[:, 3] < 256).all()
    r = ops.collisions(c1, bench.N_CODES)
    assert r["n_unique"] == s1["n_unique"] and r["n_rows"] == n - s1["n_unique"] + r["n_groups"]


def test_generate_indices_empty_and_unique_inputs(golden):
    g, m, p = load_small(golden, "small_model")
    codes, stats = G.generate_codes(m, torch.zeros(0, 96, device=DEV))
    assert codes.shape == (0, 4)
    x = T(g["x"][:1])
    codes, stats = G.generate_codes(m, x)
    assert stats["rounds"] == 0 and stats["n_unique"] == 1
    # all-identical items: one group of 64, Sinkhorn spreads them over distinct codes
    xx = T(np.repeat(g["x"][:1], 64, axis=0))
    codes, stats = G.generate_codes(m, xx)
    ref, _ = O.generate_indices(np.repeat(g["x"][:1], 64, axis=0), p, reencode=False)
    assert stats["n_unique"] == O.n_unique_codes(ref)


def test_training_step_gradients_match_torch(golden):
    """Forward through the tcgen05 kernels + autograd == plain torch fp32 model (loss and gradients)."""
    g, m, p = load_small(golden, "small_model")
    m.train()
    x = T(g["x"][:256])
    out, rq_loss, idx = m(x, use_sk=True)
    loss, rec = m.compute_loss(out, rq_loss, xs=x)
    loss.backward()
    # torch-only replica of the same forward with the same indices
    ws = [l for l in m.encoder.mlp_layers if isinstance(l, torch.nn.Linear)]
    wd = [l for l in m.decoder.mlp_layers if isinstance(l, torch.nn.Linear)]
    params = [q.detach().clone().requires_grad_(True) for q in m.parameters()]
    names = [n_ for n_, _ in m.named_parameters()]
    P = dict(zip(names, params))

    def mlp(h, prefix, lins):
        keys = sorted({int(n_.split(".")[2]) for n_ in names if n_.startswith(prefix)})
        for i, kidx in enumerate(keys):
            h = torch.nn.functional.linear(h, P[f"{prefix}.mlp_layers.{kidx}.weight"], P[f"{prefix}.mlp_layers.{kidx}.bias"])
            if i != len(keys) - 1:
                h = torch.relu(h)
        return h
    z = mlp(x, "encoder", ws)
    resid, xq, losses = z, 0, []
    for l in range(4):
        cb = P[f"rq.vq_layers.{l}.embedding.weight"]
        q = cb[idx[:, l]]
        losses.append(torch.nn.functional.mse_loss(q, resid.detach()) + 0.25 * torch.nn.functional.mse_loss(q.detach(), resid))
        xr = resid + (q - resid).detach()
        resid = resid - xr
        xq = xq + xr
    out2 = mlp(xq, "decoder", wd)
    loss2 = torch.nn.functional.mse_loss(out2, x) + torch.stack(losses).mean()
    loss2.backward()
    np.testing.assert_allclose(loss.item(), loss2.item(), rtol=1e-5)
    for (n_, a), b in zip(m.named_parameters(), params):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-3, atol=1e-7, err_msg=n_)


def test_trainer_matches_reference_losses(golden, tmp_path):
    """Reference Trainer._train_epoch loss trajectory (4 epochs, AdamW, linear warm-up, clip 1.0)."""
    import argparse
    from lcrec_b200.trainer import Trainer
    g = golden("trainer_steps")
    args = argparse.Namespace(lr=1e-3, epochs=4, batch_size=256, num_workers=0, eval_step=50, learner="AdamW",
                              lr_scheduler_type="linear", warmup_epochs=1, data_path="", weight_decay=1e-4,
                              dropout_prob=0.0, bn=False, loss_type="mse", kmeans_init=False, kmeans_iters=10,
                              sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50, device="cuda:0",
                              num_emb_list=[32] * 4, e_dim=16, quant_loss_weight=1.0, beta=0.25, layers=[64, 48],
                              save_limit=5, ckpt_dir=str(tmp_path))
    m = RQVAE(in_dim=96, num_emb_list=args.num_emb_list, e_dim=16, layers=args.layers, kmeans_init=False,
              sk_epsilons=args.sk_epsilons, sk_iters=50)
    m.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init/")})
    loader = torch.utils.data.DataLoader(torch.from_numpy(g["x"]), batch_size=256, shuffle=False)
    tr = Trainer(args, m, len(loader))
    losses = [tr._train_epoch(loader, ep) for ep in range(4)]
    np.testing.assert_allclose(np.array(losses), g["losses"], rtol=2e-3)
    coll = tr._valid_epoch(loader)
    assert abs(coll - float(g["collision_rate"])) < 0.02


@pytest.mark.gpu
def test_shared_reciprocal_division_is_ieee():
    """The last column step of the per-group Sinkhorn divides every row of a column by the same sum; the warp kernels form the
    reciprocal once per column (the sequence of the device's own double division) - its quotients must be the IEEE ones on
    everything the range check lets through: uniform operands, shares of a column sum (q <= cs, down to e^-660), near-ties."""
    from lcrec_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(11)
    n = 1 << 23
    cases = []
    a = torch.rand(n, dtype=torch.float64, device=DEV, generator=g) + 2.0 ** -60
    b = torch.rand(n, dtype=torch.float64, device=DEV, generator=g) + 2.0 ** -60
    cases.append((a, b))
    # shares: cs = q + others, others spanning 600 e-folds either way
    q = torch.exp((torch.rand(n, dtype=torch.float64, device=DEV, generator=g) - 0.5) * 600.0)
    others = q * torch.exp((torch.rand(n, dtype=torch.float64, device=DEV, generator=g) - 0.5) * 1300.0)
    cases.append((q, q + others))
    # dominated columns: cs = q (1 + k 2^-52), k small
    k = torch.randint(0, 64, (n,), device=DEV, generator=g).to(torch.float64)
    cases.append((q, q * (1.0 + k * 2.0 ** -52)))
    # wide dynamic range
    e1 = torch.exp((torch.rand(n, dtype=torch.float64, device=DEV, generator=g) - 0.5) * 1200.0)
    cases.append((a * e1, b))
    total_fallback = 0
    for x, y in cases:
        wrong, fallback = ops.ddiv_probe(x, y)
        assert wrong == 0, (wrong, fallback)
        total_fallback += fallback
    assert total_fallback < 0.2 * n * len(cases)      # the fast sequence covers the working range (fallbacks: subnormal / overflowing quotients)
