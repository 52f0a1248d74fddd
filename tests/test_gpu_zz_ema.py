"""B200 parity of the EMA-codebook variant (reference index_improve/models/vq.py) against the oracle
and the golden vectors of the unmodified reference (tests/golden/ema_kat.npz).  The EMA statistics and the codebook are
floating point, but the kernel reproduces the reference's summation order and roundings, so the bar here is BIT equality.
"""
import numpy as np
import pytest
import torch

from oracle import lcrec_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200 import ops
    from lcrec_b200.index_improve.models import RQVAE, VectorQuantizer
    DEV = torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_ema_update_kernel_matches_reference_bitwise(golden, case):
    """Teacher-forced on the reference's indices: buffers after every step == reference, bit for bit."""
    g = golden("ema_kat")
    n_e, e_dim, _, _ = (int(v) for v in g[f"c{case}_cfg"])
    cb, cs, w = T(g[f"c{case}_codebook0"]), torch.zeros(n_e, device=DEV), torch.zeros(n_e, e_dim, device=DEV)
    for s in range(int(g[f"c{case}_steps"])):
        ops.ema_update(T(g[f"c{case}_s{s}_x"]), T(g[f"c{case}_s{s}_idx"]), cs, w, cb, 0.99, 1e-5)
        if f"c{case}_s{s}_pre_reset_cs" in g:
            assert np.array_equal(cs.cpu().numpy(), g[f"c{case}_s{s}_pre_reset_cs"])
            assert np.array_equal(cb.cpu().numpy(), g[f"c{case}_s{s}_pre_reset_codebook"])
            used, unused = ops.codebook_usage(cs, 1e-5, 1e-5, want_unused=True)
            assert np.array_equal(np.nonzero(unused.cpu().numpy())[0], O.unused_codes(cs.cpu().numpy()))
            cb, cs, w = T(g[f"c{case}_s{s}_codebook"]), T(g[f"c{case}_s{s}_cs"]), T(g[f"c{case}_s{s}_w"])
        else:
            assert np.array_equal(cb.cpu().numpy(), g[f"c{case}_s{s}_codebook"])
            assert np.array_equal(cs.cpu().numpy(), g[f"c{case}_s{s}_cs"])
            assert np.array_equal(w.cpu().numpy(), g[f"c{case}_s{s}_w"])
        used, _ = ops.codebook_usage(cs, 1e-5, 1e-5)
        assert used == int(g[f"c{case}_s{s}_usage"][1])


@pytest.mark.parametrize("case", [0, 1])
def test_ema_module_forward_matches_reference(golden, case):
    """The module end to end (index selection on the kernels, Sinkhorn in case 1): same indices, loss within 1e-5,
    identical buffers, get_codebook_usage."""
    g = golden("ema_kat")
    n_e, e_dim, use_sk, _ = (int(v) for v in g[f"c{case}_cfg"])
    vq = VectorQuantizer(n_e, e_dim, beta=0.25, sk_epsilon=float(g[f"c{case}_sk_eps"]), sk_iters=50).to(DEV).train()
    vq.embedding.weight.data.copy_(T(g[f"c{case}_codebook0"]))
    for s in range(int(g[f"c{case}_steps"])):
        x = T(g[f"c{case}_s{s}_x"]).requires_grad_(True)
        x_q, loss, idx = vq(x, use_sk=bool(use_sk), use_ema=True)
        assert np.array_equal(idx.cpu().numpy(), g[f"c{case}_s{s}_idx"])
        np.testing.assert_allclose(loss.item(), g[f"c{case}_s{s}_loss"], rtol=1e-5)
        loss.backward()                                          # the in-place EMA step must not break autograd
        assert vq.embedding.weight.grad is not None and torch.isfinite(x.grad).all()
        vq.embedding.weight.grad = None
        assert np.array_equal(vq.embedding.weight.detach().cpu().numpy(), g[f"c{case}_s{s}_codebook"])
        assert np.array_equal(vq._ema_cluster_size.cpu().numpy(), g[f"c{case}_s{s}_cs"])
        assert np.array_equal(vq._ema_w.cpu().numpy(), g[f"c{case}_s{s}_w"])
        u = vq.get_codebook_usage()
        assert [u["utilization"], u["used_codes"], u["total_codes"]] == g[f"c{case}_s{s}_usage"].tolist()
    before = vq.embedding.weight.detach().clone()
    vq(x.detach(), use_sk=bool(use_sk), use_ema=False)           # use_ema=False leaves the codebook alone
    vq.eval()
    vq(x.detach(), use_sk=bool(use_sk), use_ema=True)            # and so does eval mode
    assert torch.equal(before, vq.embedding.weight.detach()) and vq.step_count == int(g[f"c{case}_steps"])


@pytest.mark.parametrize("n,k,d,skew", [(1, 4, 8, 0.0), (0, 8, 4, 0.0), (5000, 256, 32, 0.0), (9000, 16, 32, 0.8),
                                        (4097, 300, 100, 0.5), (70000, 256, 32, 0.3), (2000, 8192, 256, 0.0)])
def test_ema_update_kernel_matches_oracle_bitwise(n, k, d, skew):
    """Seeded cases incl. empty / single row, codes with more members than the kernel's list buffer (several flushes),
    e_dim beyond the CTA width, the large-codebook shape; two consecutive steps."""
    rng = np.random.default_rng(n + 31 * k + d)
    st = O.EmaState((rng.standard_normal((k, d)) * 0.05).astype(np.float32),
                    (rng.random(k) * (rng.random(k) < 0.7)).astype(np.float32),
                    (rng.standard_normal((k, d)) * 0.01).astype(np.float32))
    cb, cs, w = T(st.codebook), T(st.cluster_size), T(st.ema_w)
    for step in range(2):
        x = (rng.standard_normal((n, d)) * 0.1).astype(np.float32)
        idx = rng.integers(0, k, size=n)
        idx = np.where(rng.random(n) < skew, 3 % k, idx).astype(np.int64)
        st = O.ema_update(x, idx, st, 0.99, 1e-5)
        ops.ema_update(T(x), T(idx), cs, w, cb, 0.99, 1e-5)
        assert np.array_equal(cs.cpu().numpy(), st.cluster_size)
        assert np.array_equal(w.cpu().numpy(), st.ema_w)
        assert np.array_equal(cb.cpu().numpy(), st.codebook)
    util, used, total = O.codebook_usage(st.cluster_size)
    assert ops.codebook_usage(cs, 1e-5, 1e-5)[0] == used


def test_ema_reset_replaces_dead_codes(golden):
    """_reset_unused_codes (vq.py:79-114) on the reference's pre-reset state: min(#dead, batch) dead codes are replaced
    by a latent row + N(0, 0.01^2) noise and their statistics zeroed; live codes are untouched."""
    g = golden("ema_kat")
    s = 2
    n_e, e_dim, _, batch = (int(v) for v in g["c2_cfg"])
    vq = VectorQuantizer(n_e, e_dim, sk_epsilon=0.0, reset_interval=3).to(DEV).train()
    vq.embedding.weight.data.copy_(T(g[f"c2_s{s}_pre_reset_codebook"]))
    vq._ema_cluster_size.copy_(T(g[f"c2_s{s}_pre_reset_cs"]))
    vq._ema_w.fill_(1.0)
    x = T(g[f"c2_s{s}_x"])
    torch.manual_seed(7)
    vq._reset_unused_codes(x)
    dead = O.unused_codes(g[f"c2_s{s}_pre_reset_cs"])
    after = vq.embedding.weight.detach().cpu().numpy()
    changed = np.nonzero((after != g[f"c2_s{s}_pre_reset_codebook"]).any(axis=1))[0]
    assert len(changed) == min(len(dead), batch) and set(changed.tolist()) <= set(dead.tolist())
    assert (vq._ema_cluster_size.cpu().numpy()[changed] == 0).all() and (vq._ema_w.cpu().numpy()[changed] == 0).all()
    live = np.setdiff1d(np.arange(n_e), changed)
    assert np.array_equal(vq._ema_cluster_size.cpu().numpy()[live], g[f"c2_s{s}_pre_reset_cs"][live])
    nearest = np.abs(after[changed][:, None, :] - g[f"c2_s{s}_x"][None]).max(-1).min(-1)
    assert (nearest < 0.06).all()                                 # a latent row + noise of sigma 0.01
    vq.step_count = 2                                             # the module triggers it on the interval
    cs_before = vq._ema_cluster_size.clone()
    vq(x, use_sk=False, use_ema=True)
    assert vq.step_count == 3 and not torch.equal(cs_before, vq._ema_cluster_size)


def test_ema_rqvae_training_step_and_state_dict():
    """RQVAE of the variant: reference state_dict keys, one optimiser step with use_ema=True moves the codebooks by the
    EMA rule (a level's codebook after the step == oracle EMA of the residuals the level saw), get_indices leaves them
    alone, usage statistics per level."""
    torch.manual_seed(0)
    m = RQVAE(in_dim=96, num_emb_list=[32, 32, 32], e_dim=16, layers=[64, 48], sk_epsilons=[0.0, 0.0, 0.003],
              sk_iters=50).to(DEV).train()
    keys = set(m.state_dict().keys())
    assert {"rq.vq_layers.0._ema_cluster_size", "rq.vq_layers.2._ema_w", "rq.vq_layers.1.embedding.weight"} <= keys
    x = torch.randn(256, 96, device=DEV)
    cb0 = [q.embedding.weight.detach().cpu().numpy().copy() for q in m.rq.vq_layers]
    z = m.encoder(x).detach()
    out, rq_loss, idx = m(x, use_sk=True, use_ema=True)
    loss, _ = m.compute_loss(out, rq_loss, xs=x)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    resid = z.cpu().numpy()
    for lvl, q in enumerate(m.rq.vq_layers):
        st = O.EmaState(cb0[lvl], np.zeros(32, np.float32), np.zeros((32, 16), np.float32))
        st = O.ema_update(resid, idx[:, lvl].cpu().numpy(), st, 0.99, 1e-5)
        np.testing.assert_allclose(q.embedding.weight.detach().cpu().numpy(), st.codebook, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(q._ema_cluster_size.cpu().numpy(), st.cluster_size, rtol=0, atol=0)
        qv = cb0[lvl][idx[:, lvl].cpu().numpy()]
        resid = (resid - (resid + (qv - resid))).astype(np.float32)
    snap = [q.embedding.weight.detach().clone() for q in m.rq.vq_layers]
    m.get_indices(x)
    assert all(torch.equal(a, q.embedding.weight.detach()) for a, q in zip(snap, m.rq.vq_layers))
    stats = m.get_codebook_usage()
    assert [s["quantizer_id"] for s in stats] == [0, 1, 2] and all(0 < s["used_codes"] <= 32 for s in stats)


def test_improve_trainer_matches_reference_losses(golden, tmp_path):
    """Loss trajectory of the UNMODIFIED index_improve Trainer._train_epoch (oracle/make_golden_ema_trainer.py: 4 epochs x 4
    batches, AdamW + linear warm-up + clip 1.0, EMA step every batch, Sinkhorn on the last level) from the same initial
    state; collision rate and codebook utilisation after training."""
    import argparse
    from lcrec_b200.index_improve.trainer import Trainer
    g = golden("ema_trainer_steps")
    args = argparse.Namespace(lr=1e-3, epochs=4, batch_size=256, num_workers=0, eval_step=50, learner="AdamW",
                              lr_scheduler_type="linear", warmup_epochs=1, data_path="", weight_decay=1e-4,
                              dropout_prob=0.0, bn=False, loss_type="mse", kmeans_init=False, kmeans_iters=10,
                              sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50, device="cuda:0", num_emb_list=[32] * 4, e_dim=16,
                              quant_loss_weight=1.0, beta=0.25, layers=[64, 48], save_limit=5, ckpt_dir=str(tmp_path),
                              ema_decay=0.99, epsilon=1e-5, reset_threshold=1e-5, reset_interval=1000)
    m = RQVAE(in_dim=96, num_emb_list=args.num_emb_list, e_dim=16, layers=args.layers, kmeans_init=False,
              sk_epsilons=args.sk_epsilons, sk_iters=50)
    m.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init/")})
    loader = torch.utils.data.DataLoader(torch.from_numpy(g["x"]), batch_size=256, shuffle=False)
    tr = Trainer(args, m, len(loader))
    losses = [tr._train_epoch(loader, ep) for ep in range(4)]
    np.testing.assert_allclose(np.array(losses), g["losses"], rtol=2e-3)
    assert abs(tr._valid_epoch(loader) - float(g["collision_rate"])) < 0.02
    avg, stats = tr._get_codebook_utilization()
    assert abs(avg - float(g["avg_utilization"])) <= 0.05
    assert all(abs(s["used_codes"] - int(u)) <= 2 for s, u in zip(stats, g["used_codes"]))
    assert all(q.step_count == 16 for q in m.rq.vq_layers)
