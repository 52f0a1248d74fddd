"""CPU checks of the host-side mirrors added for the SURVEY 8(f) rows and the training path (no kernels run): interface
surface of the EMA variant (reference index_improve/), deferred host checks, optimiser selection and state layout, and
that none of the new operators has a CPU path."""
import types

import numpy as np
import pytest
import torch

from lcrec_b200 import ops
from lcrec_b200.index_improve.main import parse_args as parse_args_improve
from lcrec_b200.index_improve.models import RQVAE as RQVAEImprove
from lcrec_b200.index_improve.models import ResidualVectorQuantizer as RQImprove
from lcrec_b200.index_improve.models import VectorQuantizer as VQImprove
from lcrec_b200.index_improve.trainer import Trainer as TrainerImprove
from lcrec_b200.models import layers as L
from lcrec_b200.optim import FusedAdam
from lcrec_b200.trainer import Trainer

# state_dict keys of the unmodified reference module (index_improve/models/rqvae.py built with two levels, one hidden layer)
REF_IMPROVE_KEYS = [
    "encoder.mlp_layers.1.weight", "encoder.mlp_layers.1.bias", "encoder.mlp_layers.4.weight", "encoder.mlp_layers.4.bias",
    "rq.vq_layers.0._ema_cluster_size", "rq.vq_layers.0._ema_w", "rq.vq_layers.0.embedding.weight",
    "rq.vq_layers.1._ema_cluster_size", "rq.vq_layers.1._ema_w", "rq.vq_layers.1.embedding.weight",
    "decoder.mlp_layers.1.weight", "decoder.mlp_layers.1.bias", "decoder.mlp_layers.4.weight", "decoder.mlp_layers.4.bias"]


def test_improve_variant_surface():
    m = RQVAEImprove(in_dim=64, num_emb_list=[8, 8], e_dim=4, layers=[16], sk_epsilons=[0.0, 0.003], ema_decay=0.9,
                     epsilon=1e-4, reset_threshold=1e-3, reset_interval=7)
    assert list(m.state_dict().keys()) == REF_IMPROVE_KEYS
    assert isinstance(m.rq, RQImprove) and all(isinstance(q, VQImprove) for q in m.rq.vq_layers)
    q = m.rq.vq_layers[1]
    assert (q.ema_decay, q.epsilon, q.reset_threshold, q.reset_interval, q.step_count) == (0.9, 1e-4, 1e-3, 7, 0)
    assert q._ema_cluster_size.shape == (8,) and q._ema_w.shape == (8, 4) and float(q._ema_w.abs().sum()) == 0.0
    assert q.sk_epsilon == 0.003 and q.initted
    for name in ("forward", "get_indices", "compute_loss", "get_codebook_usage"):
        assert callable(getattr(m, name))
    import inspect
    assert list(inspect.signature(m.forward).parameters) == ["x", "use_sk", "use_ema"]
    assert list(inspect.signature(q.forward).parameters) == ["x", "use_sk", "use_ema"]
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        q.get_codebook_usage()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        q.train()(torch.zeros(5, 4))


def test_improve_cli_flags():
    a = parse_args_improve([])
    assert a.sk_epsilons is None and a.ema_decay == 0.99 and a.epsilon == 1e-5 and a.reset_threshold == 1e-5
    assert a.reset_interval == 1000 and a.num_emb_list == [256, 256, 256] and a.kmeans_init is True
    b = parse_args_improve(["--ema_decay", "0.95", "--reset_interval", "10", "--sk_epsilons", "0", "0", "0.003", "--bn", "False"])
    assert b.ema_decay == 0.95 and b.reset_interval == 10 and b.sk_epsilons == [0.0, 0.0, 0.003] and b.bn is True


def _trainer_args(tmp_path, learner="AdamW"):
    return types.SimpleNamespace(lr=1e-3, learner=learner, lr_scheduler_type="linear", weight_decay=1e-4, epochs=4,
                                 warmup_epochs=1, save_limit=2, eval_step=2, device="cpu", ckpt_dir=str(tmp_path))


def test_trainer_optimizer_selection_and_improve_log(tmp_path):
    """On a CPU device the Trainer keeps torch.optim (FusedAdam is CUDA only); the variant's trainer passes use_ema and
    formats the utilisation lines of index_improve/trainer.py:246-253."""
    m = RQVAEImprove(in_dim=64, num_emb_list=[8, 8], e_dim=4, layers=[16], sk_epsilons=[0.0, 0.0])
    t = TrainerImprove(_trainer_args(tmp_path), m, data_num=3)
    assert isinstance(t.optimizer, torch.optim.AdamW) and not isinstance(t.optimizer, FusedAdam)
    assert isinstance(Trainer(_trainer_args(tmp_path, "sgd"), m, 3).optimizer, torch.optim.SGD)
    seen = {}
    m.forward = lambda data, **kw: seen.update(kw) or ("out", "loss", "idx")
    assert t._model_forward(torch.zeros(1)) == ("out", "loss", "idx") and seen == {"use_ema": True}
    m.get_codebook_usage = lambda: [{"utilization": 0.5, "used_codes": 4, "total_codes": 8, "quantizer_id": 0},
                                    {"utilization": 0.25, "used_codes": 2, "total_codes": 8, "quantizer_id": 1}]
    line = t._generate_valid_output(3, 1.5, 0.125)
    assert "0.1250" in line and "0.3750" in line and "Quantizer 1: 0.2500 (2/8)" in line and line.endswith("]")
    base = Trainer(_trainer_args(tmp_path), m, 3)._generate_valid_output(3, 1.5, 0.125)
    assert "0.125000]" in base and "Quantizer" not in base


def test_fused_adam_state_layout_matches_torch():
    p = [torch.nn.Parameter(torch.zeros(3))]
    ours, ref = FusedAdam(p, lr=2e-3, weight_decay=1e-4), torch.optim.AdamW(p, lr=2e-3, weight_decay=1e-4)
    ga, gb = ours.state_dict()["param_groups"][0], ref.state_dict()["param_groups"][0]
    assert set(ga.keys()) == set(gb.keys())
    assert all(ga[k] == gb[k] for k in ("lr", "betas", "eps", "weight_decay", "decoupled_weight_decay", "params"))
    assert FusedAdam(p, decoupled=False).param_groups[0]["decoupled_weight_decay"] is False
    p[0].grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ours.step()
    sched = torch.optim.lr_scheduler.LambdaLR(ours, lambda i: 0.5)
    assert ours.param_groups[0]["lr"] == 1e-3 and sched.get_last_lr() == [1e-3]


def test_deferred_checks_queue_and_flush(capsys):
    bad, ok, nan = torch.tensor([1], dtype=torch.int32), torch.tensor([0], dtype=torch.int32), torch.tensor([1], dtype=torch.int32)
    ops.check_later("amplitude", ok)
    with pytest.raises(AssertionError, match="amplitude > 0"):
        ops.check_later("amplitude", bad)                          # immediate, like the reference's assert (vq.py:59)
    with pytest.raises(AssertionError):
        with ops.defer_checks():
            ops.check_later("amplitude", bad)
            ops.check_later("sinkhorn", nan)
            assert len(ops._PENDING) == 2
    assert ops._PENDING == []
    with ops.defer_checks():
        with ops.defer_checks():                                    # nested: evaluated when the outermost block ends
            ops.check_later("sinkhorn", nan)
        assert len(ops._PENDING) == 1
    assert ops._PENDING == [] and "Sinkhorn Algorithm returns nan/inf values." in capsys.readouterr().out
    with pytest.raises(ValueError):
        with ops.defer_checks():
            ops.check_later("amplitude", bad)
            raise ValueError("step failed first")
    assert ops._PENDING == []                                      # nothing leaks into the next step
    with pytest.raises(RuntimeError, match="peer rank"):
        ops.check_later("sinkhorn", torch.tensor([8], dtype=torch.int32))


def test_new_operators_have_no_cpu_path(tmp_path):
    x = torch.zeros(4, 8)
    for call in (lambda: ops.ema_update(x, torch.zeros(4, dtype=torch.int64), torch.zeros(2), torch.zeros(2, 8), torch.zeros(2, 8), 0.99, 1e-5),
                 lambda: ops.codebook_usage(torch.zeros(4), 1e-5, 1e-5),
                 lambda: ops.masked_mean_pool(torch.zeros(2, 3, 8), torch.ones(2, 3, dtype=torch.int64)),
                 lambda: ops.kmeans_center(x),
                 lambda: ops.kmeans_lloyd(x, x[:2], 5, 1e-4),
                 lambda: ops.rq_train_forward(x, torch.zeros(4, 1, dtype=torch.int64), [torch.zeros(2, 8)]),
                 lambda: ops.rq_train_backward(torch.zeros(1, 4, 8), torch.zeros(1, 4, dtype=torch.int64), [2], None, None, 0.25)):
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            call()
    old = L.KMEANS_BACKEND
    try:
        L.KMEANS_BACKEND = "device"
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            L.kmeans(x, 2, 5)
        L.KMEANS_BACKEND = "gpu?"
        with pytest.raises(ValueError):
            L.kmeans(x, 2, 5)
        L.KMEANS_BACKEND = "sklearn"                                # the reference's own call works anywhere
        np.random.seed(0)
        c = L.kmeans(torch.randn(64, 8), 4, 5)
        assert c.shape == (4, 8) and c.dtype == torch.float32
    finally:
        L.KMEANS_BACKEND = old
    from lcrec_b200.text_emb import generate_item_embedding
    args = types.SimpleNamespace(root=str(tmp_path), dataset="Toy", plm_name="p", max_sent_len=8, device="cpu")
    tok = lambda *a, **k: _Enc()                                    # noqa: E731

    class _Enc(dict):
        def __init__(self):
            super().__init__(input_ids=torch.ones(1, 3, dtype=torch.int64), attention_mask=torch.ones(1, 3, dtype=torch.int64))
        to = lambda self, device: self                              # noqa: E731
        input_ids = property(lambda self: self["input_ids"])
        attention_mask = property(lambda self: self["attention_mask"])
    model = lambda input_ids, attention_mask: types.SimpleNamespace(last_hidden_state=torch.zeros(1, 3, 8))   # noqa: E731
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        generate_item_embedding(args, [[0, ["a b", "c"]]], tok, model)


def test_rq_training_backward_formulas_match_autograd():
    """The analytic gradients lcrec_rq_train_backward implements (include/lcrec_b200.h), evaluated here with plain torch on
    the CPU, against autograd through the reference's own per-level expressions (vq.py:87-99, rq.py:39-56) for fixed codes:
    d x_q / d z = I, only level 0's commitment term reaches z, codebooks receive the codebook-loss term."""
    torch.manual_seed(0)
    n, d, ks, beta = 50, 6, [7, 5, 9], 0.25
    z = torch.randn(n, d, dtype=torch.float64, requires_grad=True)
    cbs = [torch.randn(k, d, dtype=torch.float64, requires_grad=True) for k in ks]
    codes = torch.stack([torch.randint(0, k, (n,)) for k in ks], dim=1)
    w = torch.randn(n, d, dtype=torch.float64)
    residual, x_q, losses, diffs = z, 0, [], []
    for l, cb in enumerate(cbs):
        q = cb[codes[:, l]]
        diffs.append((q - residual).detach())
        loss = torch.nn.functional.mse_loss(q, residual.detach()) + beta * torch.nn.functional.mse_loss(q.detach(), residual)
        x_res = residual + (q - residual).detach()
        residual = residual - x_res
        x_q = x_q + x_res
        losses.append(loss)
    mean_loss = torch.stack(losses).mean()
    g_loss = 3.0
    ((x_q * w).sum() + g_loss * mean_loss).backward()
    per_elem = 2.0 / (n * d) / len(ks)
    np.testing.assert_allclose(z.grad.numpy(), (w - g_loss * beta * per_elem * diffs[0]).numpy(), rtol=1e-12, atol=1e-15)
    for l, cb in enumerate(cbs):
        want = torch.zeros_like(cb).index_add_(0, codes[:, l], diffs[l]) * (g_loss * per_elem)
        np.testing.assert_allclose(cb.grad.numpy(), want.detach().numpy(), rtol=1e-12, atol=1e-15)


def test_fused_adam_marks_optimizer_step_for_schedulers(recwarn):
    p = [torch.nn.Parameter(torch.zeros(3))]
    opt = FusedAdam(p, lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda i: 1.0)
    opt.clip_and_step(1.0)                       # no gradients: nothing to launch, but the call is recorded
    sched.step()
    assert not [w for w in recwarn.list if "lr_scheduler.step()" in str(w.message)]


def test_producer_host_loop_against_reference_npy(golden, tmp_path, monkeypatch):
    """Host logic of lcrec_b200.text_emb.generate_item_embedding (item ordering, per-field accumulate / divide, batching,
    word drop draws, file contract) with the pooling operator replaced by the ORACLE (tests only - the product operator has
    no CPU path): reproduces the .npy of the unmodified reference producer, also when several items are pooled per call."""
    import os
    import random
    from oracle import lcrec_oracle as O
    from lcrec_b200 import text_emb
    g = golden("pool_kat")
    n, nf = int(g["n_items"]), int(g["n_fields"])

    def oracle_pool(h, mask, out=None, accumulate=False, divide_by=0.0):
        r = torch.from_numpy(O.masked_mean_pool(h.numpy(), mask.numpy()))
        if accumulate:
            r = out + r
        if divide_by > 0:
            r = r / divide_by
        out.copy_(r)
        return out
    monkeypatch.setattr(text_emb.ops, "masked_mean_pool", oracle_pool)
    E, P = torch.from_numpy(g["E"]), torch.from_numpy(g["P"])
    ids_of = {}                                                     # (item, field) -> recorded token row
    for i in range(n):
        for f in range(nf):
            ids, mask = g[f"ids_{nf * i + f}"], g[f"mask_{nf * i + f}"]
            ids_of[f"item{i}field{f}"] = ids[0][mask[0] == 1]

    class Enc(dict):
        to = lambda self, device: self                              # noqa: E731
        input_ids = property(lambda self: self["input_ids"])
        attention_mask = property(lambda self: self["attention_mask"])

    def tokenizer(sentences, max_length, truncation, return_tensors, padding):
        rows = [ids_of[s.split(" ")[0]] for s in sentences]         # first word names the recorded sequence
        width = (max(len(r) for r in rows) + 7) // 8 * 8
        ids = torch.zeros(len(rows), width, dtype=torch.int64)
        mask = torch.zeros(len(rows), width, dtype=torch.int64)
        for j, r in enumerate(rows):
            ids[j, : len(r)] = torch.from_numpy(np.asarray(r))
            mask[j, : len(r)] = 1
        return Enc(input_ids=ids, attention_mask=mask)

    def model(input_ids, attention_mask):
        h = E[input_ids] + P[: input_ids.shape[1]][None] + (1 - attention_mask).unsqueeze(-1) * 1e3
        return types.SimpleNamespace(last_hidden_state=h)

    items = [[i, [f"item{i}field0 x y", f"item{i}field1 z"]] for i in reversed(range(n))]
    scale = np.abs(g["emb"]).max(axis=1, keepdims=True)
    for bs in (1, 5):
        args = types.SimpleNamespace(root=str(tmp_path), dataset=f"Toy{bs}", plm_name="standin", max_sent_len=24, device="cpu")
        emb = text_emb.generate_item_embedding(args, items, tokenizer, model, word_drop_ratio=-1, batch_size=bs)
        saved = np.load(os.path.join(str(tmp_path), f"Toy{bs}.emb-standin-td.npy"))
        assert saved.shape == g["emb"].shape and np.array_equal(saved, emb.numpy())
        assert (np.abs(saved - g["emb"]) <= 1e-6 * scale).all()
    # word drop consumes Python's RNG once per word, in order (amazon_text_emb.py:73-85)
    random.seed(3)
    expect = [random.random() for _ in range(5)]
    random.seed(3)
    seen = []
    tok2 = lambda sentences, **kw: seen.append(list(sentences)) or tokenizer([f"item0field{len(seen) - 1}"], **kw)   # noqa: E731
    args = types.SimpleNamespace(root=str(tmp_path), dataset="ToyDrop", plm_name="standin", max_sent_len=24, device="cpu")
    text_emb.generate_item_embedding(args, [[0, ["a b c", "d e"]]], tok2, model, word_drop_ratio=0.5, save=False)
    kept = [" ".join(w for w, r in zip("a b c".split(), expect[:3]) if r > 0.5),
            " ".join(w for w, r in zip("d e".split(), expect[3:]) if r > 0.5)]
    assert seen == [[kept[0]], [kept[1]]]


def test_kmeanspp_random_numbers_follow_sklearn():
    """Host half of the device seeding: the first centre and the per-centre uniforms are what sklearn's kmeans_plusplus
    consumes from numpy's global RNG (same first index, same RNG state afterwards), independent of the data."""
    from sklearn.cluster import kmeans_plusplus
    from lcrec_b200.models.layers import kmeanspp_random_numbers
    from oracle import lcrec_oracle as O
    rng = np.random.default_rng(0)
    x = rng.standard_normal((700, 12)).astype(np.float32)
    xc = x - x.mean(axis=0)
    np.random.seed(11)
    _, want = kmeans_plusplus(xc, 40)
    after = np.random.random_sample()
    np.random.seed(11)
    first, draws = kmeanspp_random_numbers(700, 40)
    assert np.random.random_sample() == after and first == want[0] and draws.shape == (39, 2 + int(np.log(40)))
    # the oracle's restatement, fed with these numbers, reproduces every seed
    p = np.ones(700, dtype=np.float64) / 700.0
    cdf = p.cumsum(); cdf /= cdf[-1]
    np.random.seed(11)
    u0, dr = O.kmeanspp_draws(np.random.mtrand._rand, 40)
    assert int(cdf.searchsorted(u0, side="right")) == first and all(np.array_equal(a, b) for a, b in zip(dr, draws))
    assert np.array_equal(O.kmeanspp_predrawn(xc, 40, u0, dr), want)


def test_level_chain_and_factories_with_stub_levels():
    """The shared per-level loop (rq.py:45-55) on stand-in level modules: residual / x_q bookkeeping, mean of the level
    losses, stacked codes, keyword pass-through (use_ema) - and the class-level factories the EMA variant overrides."""
    from lcrec_b200.models import RQVAE, ResidualVectorQuantizer, VectorQuantizer

    class Stub(torch.nn.Module):
        def __init__(self, k):
            super().__init__()
            self.k, self.seen = k, None

        def forward(self, r, use_sk=True, **kw):
            self.seen = dict(use_sk=use_sk, **kw)
            return r * 0.5, (r.detach() ** 2).mean() * self.k, torch.full((r.shape[0],), self.k)
    rq = ResidualVectorQuantizer([4, 4], 3, [0.0, 0.0])
    assert ResidualVectorQuantizer._LEVEL is VectorQuantizer and RQVAE._RQ is ResidualVectorQuantizer
    assert RQImprove._LEVEL is VQImprove and RQVAEImprove._RQ is RQImprove
    rq.vq_layers = torch.nn.ModuleList([Stub(1), Stub(2)])
    x = torch.ones(2, 3)
    x_q, loss, idx = rq._forward_levels(x, use_sk=False, use_ema=True)
    assert torch.allclose(x_q, torch.full((2, 3), 0.75)) and abs(float(loss) - 0.75) < 1e-7
    assert idx.tolist() == [[1, 2], [1, 2]] and rq.vq_layers[1].seen == {"use_sk": False, "use_ema": True}
    m = RQVAE(in_dim=10, num_emb_list=[4], e_dim=3, layers=[5], sk_epsilons=[0.0], loss_type="l1")
    out, xs = torch.zeros(2, 10), torch.ones(2, 10)
    total, recon = m.compute_loss(out, torch.tensor(0.5), xs=xs)
    assert float(recon) == 1.0 and float(total) == 1.5
    m.loss_type = "huber"
    with pytest.raises(ValueError, match="incompatible loss type"):
        m.compute_loss(out, torch.tensor(0.5), xs=xs)
