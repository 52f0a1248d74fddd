"""Pin the numpy oracle against outputs of the unmodified reference (tests/golden, produced by
oracle/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from oracle import lcrec_oracle as O
from tests.conftest import state_dict_of
from lcrec_b200.synth import seeded_weights, synth_items


def test_sinkhorn_matches_reference(golden):
    g = golden("sinkhorn_kat")
    for ci, (n, k, eps, iters) in enumerate(g["meta"]):
        d32 = g[f"d_{ci}"]
        dc = O.center_distance_for_constraint(d32)
        np.testing.assert_allclose(dc, g[f"dc_{ci}"], rtol=0, atol=2e-7)
        # teacher-forced: reference centred distances in, exact argmax out
        q = O.sinkhorn_algorithm(g[f"dc_{ci}"].astype(np.float64), float(eps), int(iters))
        arg = O._argmax_first_nan_wins(q)
        assert (arg == g[f"arg_{ci}"]).all(), f"case {ci}"
        qr = g[f"q_{ci}"]
        np.testing.assert_allclose(q[: qr.shape[0]], qr, rtol=1e-9, atol=1e-300)


def test_sinkhorn_exact_ties_present(golden):
    """F3: duplicate rows give bit-exact ties that the lowest index must win."""
    g = golden("sinkhorn_kat")
    q = g["q_7"]
    assert (q[0] == q[1]).all() and (q[1] == q[2]).all()
    assert g["arg_7"][0] == g["arg_7"][1] == g["arg_7"][2]


@pytest.mark.parametrize("name", ["small_model", "bn_model"])
def test_model_forward_matches_reference(golden, name):
    g = golden(name)
    p = O.params_from_state_dict(state_dict_of(g), g["sk_epsilons"].tolist(), int(g["sk_iters"]))
    z = O.mlp_forward(g["x"], p.encoder)
    np.testing.assert_allclose(z, g["latents"], rtol=2e-5, atol=2e-6)
    # teacher-forced RQ on reference latents
    xq, loss, idx = O.rq_forward(g["latents"], p, use_sk=False)
    near, hard = O.classify_code_mismatches(g["latents"], p, g["codes_argmin_full"])
    assert hard == 0 and near <= 2
    assert (idx != g["codes_argmin_full"]).any(axis=1).sum() <= 2
    np.testing.assert_allclose(xq, g["rq_xq"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(loss, g["rq_loss"], rtol=1e-5)
    out, rq_loss, fidx, total, recon = O.rqvae_forward(g["x"][:512], p, use_sk=True)
    assert (fidx != g["fwd_idx"]).any(axis=1).mean() < 0.01
    np.testing.assert_allclose(recon, g["fwd_recon"], rtol=1e-4)
    np.testing.assert_allclose(total, g["fwd_total"], rtol=1e-4)


@pytest.mark.parametrize("name", ["small_model", "bn_model"])
def test_generate_indices_matches_reference_script(golden, name):
    g = golden(name)
    p = O.params_from_state_dict(state_dict_of(g), g["sk_epsilons"].tolist(), int(g["sk_iters"]))
    codes, trace = O.generate_indices(g["x"], p, batch_size=64, max_rounds=20, reencode=True)
    assert (trace.codes_pass0 != g["codes_pass0"]).any(axis=1).sum() <= 2
    # teacher-forced per round: start every round from the reference's table
    tables = [g["codes_pass0"]] + list(g["rounds"])
    p_sk = O.RqvaeParams(encoder=p.encoder, codebooks=p.codebooks, sk_epsilons=O.generation_epsilons(p),
                         sk_iters=p.sk_iters, beta=p.beta)
    bad_rows = tot_rows = 0
    for t in range(min(4, len(tables) - 1)):
        cur = tables[t].copy()
        for grp in O.collision_groups(cur):
            cur[grp] = O.rq_forward(O.mlp_forward(g["x"][grp], p.encoder), p_sk, use_sk=True)[2]
            tot_rows += len(grp)
        bad_rows += int((cur != tables[t + 1]).any(axis=1).sum())
    assert bad_rows <= max(2, tot_rows // 500), (bad_rows, tot_rows)
    # end to end: same collision statistics as the script
    ref_final = g["script_codes_final"]
    assert abs(O.collision_rate(codes) - O.collision_rate(ref_final)) < 5e-3
    assert (codes != ref_final).any(axis=1).mean() < 0.05
    assert O.index_json(ref_final).encode() == g["script_json"].tobytes()


def test_fullshape_matches_reference(golden):
    g = golden("fullshape")
    dims = g["dims"].tolist()
    ws, bs, cbs = seeded_weights(dims, [256] * 4, 32, seed=int(g["seed_w"]))
    x = synth_items(int(g["n"]), dims[0], n_parents=int(g["n"]) // 8, seed=int(g["seed_x"]))
    p = O.RqvaeParams(encoder=O.MlpParams(ws, bs), codebooks=cbs, sk_epsilons=[0, 0, 0, 0.003])
    z = O.mlp_forward(x, p.encoder)
    np.testing.assert_allclose(z, g["latents"], rtol=1e-4, atol=1e-6)
    near, hard = O.classify_code_mismatches(g["latents"], p, g["codes"])
    assert hard == 0
    codes = O.rq_forward(z, p, use_sk=False)[2]
    assert (codes != g["codes"]).any(axis=1).sum() <= 2


def test_collision_bookkeeping():
    codes = np.array([[1, 2, 3, 4], [0, 0, 0, 0], [1, 2, 3, 4], [5, 5, 5, 5], [0, 0, 0, 0], [1, 2, 3, 4]])
    assert O.collision_groups(codes) == [[0, 2, 5], [1, 4]]
    assert O.n_unique_codes(codes) == 3
    assert O.max_conflicts(codes) == 3
    assert abs(O.collision_rate(codes) - 0.5) < 1e-12
    assert json.loads(O.index_json(codes[:1])) == {"0": ["<a_1>", "<b_2>", "<c_3>", "<d_4>"]}


# ------------------------------------------------------------------ f3: index_improve EMA quantiser
@pytest.mark.parametrize("case", [0, 1, 2])
def test_ema_quantiser_matches_reference(golden, case):
    """Oracle vs the unmodified index_improve VectorQuantizer in training mode (oracle/make_golden_ema.py): indices,
    loss, and BIT-identical codebook / _ema_cluster_size / _ema_w after every step; usage counts; dead-code set."""
    g = golden("ema_kat")
    n_e, e_dim, use_sk, _ = (int(v) for v in g[f"c{case}_cfg"])
    eps = float(g[f"c{case}_sk_eps"])
    st = O.EmaState(g[f"c{case}_codebook0"].copy(), np.zeros(n_e, np.float32), np.zeros((n_e, e_dim), np.float32))
    resets = 0
    for s in range(int(g[f"c{case}_steps"])):
        x = g[f"c{case}_s{s}_x"]
        _, loss, idx, st = O.vq_forward_ema(x, st, bool(use_sk), eps, 50, 0.25)
        assert (idx == g[f"c{case}_s{s}_idx"]).all()
        np.testing.assert_allclose(loss, g[f"c{case}_s{s}_loss"], rtol=1e-5)
        if f"c{case}_s{s}_pre_reset_cs" in g:          # this step crossed reset_interval (RNG-driven; teacher-forced)
            resets += 1
            assert np.array_equal(st.cluster_size, g[f"c{case}_s{s}_pre_reset_cs"])
            assert np.array_equal(st.codebook, g[f"c{case}_s{s}_pre_reset_codebook"])
            dead = set(O.unused_codes(st.cluster_size).tolist())
            after = g[f"c{case}_s{s}_codebook"]
            changed = np.nonzero((after != st.codebook).any(axis=1))[0]
            assert len(changed) == min(len(dead), x.shape[0]) and set(changed.tolist()) <= dead
            assert (g[f"c{case}_s{s}_cs"][changed] == 0).all() and (g[f"c{case}_s{s}_w"][changed] == 0).all()
            st = O.EmaState(after.copy(), g[f"c{case}_s{s}_cs"].copy(), g[f"c{case}_s{s}_w"].copy(), st.step_count)
        else:
            assert np.array_equal(st.codebook, g[f"c{case}_s{s}_codebook"])
            assert np.array_equal(st.cluster_size, g[f"c{case}_s{s}_cs"])
            assert np.array_equal(st.ema_w, g[f"c{case}_s{s}_w"])
        util, used, total = O.codebook_usage(st.cluster_size)
        assert [util, used, total] == g[f"c{case}_s{s}_usage"].tolist()
    assert resets == (1 if case == 2 else 0)


# ------------------------------------------------------------------ f4: embedding producer hand-off
def _standin_hidden(g, j):
    ids, mask = g[f"ids_{j}"], g[f"mask_{j}"]
    h = g["E"][ids] + g["P"][: ids.shape[1]][None]
    return (h + (1 - mask)[..., None] * np.float32(1e3)).astype(np.float32), mask


def test_masked_mean_pool_matches_reference_producer(golden):
    """Oracle vs the .npy the unmodified generate_item_embedding wrote (oracle/make_golden_pool.py).  Floating point:
    torch's CPU sum over the sequence is a cascade, the oracle adds in position order - bar 1e-6 of the row scale
    (measured 1.6e-7; bit-equal for sequences up to 16 positions)."""
    g = golden("pool_kat")
    nf = int(g["n_fields"])
    for i in range(int(g["n_items"])):
        hs, ms = zip(*[_standin_hidden(g, nf * i + f) for f in range(nf)])
        assert any((m == 0).any() for m in ms) or i > 0
        e = O.item_embedding(hs, ms)[0]
        assert np.abs(e - g["emb"][i]).max() <= 1e-6 * np.abs(g["emb"][i]).max()


# ------------------------------------------------------------------ f1: k-means++ seeding with pre-drawn random numbers
@pytest.mark.parametrize("n,d,k,seed", [(1024, 32, 256, 0), (600, 100, 20, 1), (2000, 16, 64, 2)])
def test_kmeanspp_with_predrawn_numbers_equals_sklearn(n, d, k, seed):
    """The oracle's restatement of scikit-learn's k-means++ (the third-party routine behind layers.py:69-82, pinned 1.9.0)
    picks the same seeds when it is handed the random numbers up front, and leaves numpy's global RNG in the same state:
    the seeding can move to the device without changing what the reference's KMeans.fit would draw."""
    from sklearn.cluster import kmeans_plusplus
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, d)) * 0.5 + 1.0).astype(np.float32)
    xc = x - x.mean(axis=0)
    np.random.seed(2024 + seed)
    _, want = kmeans_plusplus(xc, k)
    state_after = np.random.random_sample()
    np.random.seed(2024 + seed)
    u0, draws = O.kmeanspp_draws(np.random.mtrand._rand, k)
    assert np.random.random_sample() == state_after
    got = O.kmeanspp_predrawn(xc, k, u0, draws)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ a11: clip + Adam / AdamW
@pytest.mark.parametrize("decoupled,wd", [(True, 1e-4), (False, 1e-2), (True, 0.0)])
def test_adam_clip_step_matches_torch_optim(decoupled, wd):
    """The oracle's restatement of clip_grad_norm_ + torch.optim.AdamW / Adam (the third-party routines behind
    trainer.py:49-81, :117-118) against torch itself on the CPU, four steps."""
    import torch
    rng = np.random.default_rng(0)
    shapes = [(37, 5), (64,), (1, 1), (300, 7)]
    p0 = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    tp = [torch.nn.Parameter(torch.from_numpy(x.copy())) for x in p0]
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(tp, lr=1e-3, weight_decay=wd)
    p, m, v = [x.copy() for x in p0], [np.zeros_like(x) for x in p0], [np.zeros_like(x) for x in p0]
    for step in range(1, 5):
        grads = [(rng.standard_normal(s) * (3.0 if step % 2 else 0.01)).astype(np.float32) for s in shapes]
        for t, gr in zip(tp, grads):
            t.grad = torch.from_numpy(gr.copy())
        want_norm = float(torch.nn.utils.clip_grad_norm_(tp, 1.0))
        opt.step()
        p, gc, m, v, norm = O.adam_clip_step(p, grads, m, v, step, 1e-3, weight_decay=wd, decoupled=decoupled, max_norm=1.0)
        np.testing.assert_allclose(norm, want_norm, rtol=1e-6)
        for a, t, c in zip(p, tp, gc):
            np.testing.assert_allclose(c, t.grad.numpy(), rtol=1e-6, atol=0)
            np.testing.assert_allclose(a, t.detach().numpy(), rtol=1e-5, atol=2e-6)
        for a, t in zip(m, tp):
            np.testing.assert_allclose(a, opt.state[t]["exp_avg"].numpy(), rtol=1e-5, atol=1e-8)
