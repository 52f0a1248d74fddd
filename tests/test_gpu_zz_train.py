"""B200 checks of the fused training-side residual quantiser (lcrec_rq_train_forward / _backward; reference
index/models/rq.py:39-56 over vq.py:87-99): values and gradients against the per-level autograd path of the same module
(which tests/test_gpu_parity.py pins to a plain torch replica and to the reference's loss trajectory), and the deferred
host checks of the training loop."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200 import ops
    from lcrec_b200.models import RQVAE, ResidualVectorQuantizer
    from lcrec_b200.models import rq as RQ
    DEV = torch.device("cuda:0")


def _run(rq, z0, use_sk, fused, w_xq):
    RQ.FUSED_TRAIN = fused
    try:
        for p in rq.parameters():
            p.grad = None
        z = z0.clone().requires_grad_(True)
        x_q, loss, idx = rq(z, use_sk=use_sk)
        ((x_q * w_xq).sum() + 3.0 * loss).backward()
        return (x_q.detach(), loss.detach(), idx, z.grad.clone(),
                [q.embedding.weight.grad.clone() for q in rq.vq_layers])
    finally:
        RQ.FUSED_TRAIN = True


@pytest.mark.parametrize("n,d,ks,eps", [(1024, 32, [256] * 4, [0.0, 0.0, 0.0, 0.003]), (300, 16, [32, 32, 32], [0.0, 0.0, 0.0]),
                                        (257, 48, [100, 50], [0.0, 0.003]), (64, 16, [32, 32, 32], [0.0, 0.003, 0.0]),
                                        (1, 32, [256], [0.0])])
def test_fused_rq_training_equals_per_level_autograd(n, d, ks, eps):
    torch.manual_seed(n + d)
    rq = ResidualVectorQuantizer(ks, d, eps, beta=0.25, sk_iters=50).to(DEV).train()
    for q in rq.vq_layers:
        q.embedding.weight.data.normal_(0, 0.3)
    z0 = torch.randn(n, d, device=DEV) * 0.5
    w = torch.randn(n, d, device=DEV)
    a = _run(rq, z0, True, True, w)
    b = _run(rq, z0, True, False, w)
    assert torch.equal(a[2], b[2])                                   # same codes
    assert torch.equal(a[0], b[0])                                   # x_q bit-identical (same operation order)
    np.testing.assert_allclose(a[1].item(), b[1].item(), rtol=2e-6)  # loss: fp64 vs fp32 tree sum of the squares
    np.testing.assert_allclose(a[3].cpu().numpy(), b[3].cpu().numpy(), rtol=1e-5, atol=1e-9)
    for ga, gb in zip(a[4], b[4]):
        np.testing.assert_allclose(ga.cpu().numpy(), gb.cpu().numpy(), rtol=1e-4, atol=1e-9)
    # deterministic: the fused path repeats bit for bit (no atomics in the codebook gradient)
    c = _run(rq, z0, True, True, w)
    assert torch.equal(a[3], c[3]) and all(torch.equal(x, y) for x, y in zip(a[4], c[4]))


def test_fused_rq_training_inside_model_step():
    """Whole model: one optimiser step with the fused node == one with the per-level path (same weights afterwards to
    fp32 rounding), eval / no-grad calls keep using the inference kernels."""
    def make():
        torch.manual_seed(3)
        return RQVAE(in_dim=96, num_emb_list=[32] * 4, e_dim=16, layers=[64, 48], sk_epsilons=[0.0, 0.0, 0.0, 0.003],
                     sk_iters=50).to(DEV).train()
    x = torch.randn(256, 96, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    outs = []
    for fused in (True, False):
        RQ.FUSED_TRAIN = fused
        try:
            m = make()
            opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
            with ops.defer_checks():
                out, rq_loss, idx = m(x, use_sk=True)
                loss, _ = m.compute_loss(out, rq_loss, xs=x)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                opt.step()
            outs.append((loss.item(), idx.clone(), [p.detach().clone() for p in m.parameters()]))
        finally:
            RQ.FUSED_TRAIN = True
    assert torch.equal(outs[0][1], outs[1][1])
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-6)
    for pa, pb in zip(outs[0][2], outs[1][2]):
        np.testing.assert_allclose(pa.cpu().numpy(), pb.cpu().numpy(), rtol=1e-4, atol=2e-6)
    m.eval()
    with torch.no_grad():
        assert m(x, use_sk=False)[2].shape == (256, 4)


def test_deferred_checks_raise_at_flush():
    """vq.py:59 `assert amplitude > 0` and the NaN print are host reads; under defer_checks they fire at flush time."""
    bad = torch.full((2, 4), float("nan"), device=DEV)
    with pytest.raises(AssertionError):
        ops.center_distances(bad)                                   # default: where the reference asserts
    with pytest.raises(AssertionError):
        with ops.defer_checks():
            ops.center_distances(bad)                               # queued ...
            assert len(ops._PENDING) == 1
        # ... and raised when the block ends
    assert len(ops._PENDING) == 0
    with ops.defer_checks():
        ops.center_distances(torch.rand(4, 8, device=DEV))
    assert len(ops._PENDING) == 0
