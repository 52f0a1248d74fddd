"""B200 checks of the fused clip + Adam / AdamW step (lcrec_adam_clip_step; reference index/trainer.py:49-81, :117-119 =
torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW / Adam): parameters, moments, clipped gradients and the total norm
against torch's own optimiser over several steps (fp32 elementwise arithmetic: bar 1e-5 relative; the absolute floors cover
the handful of elements in 8 M where `g * coef + wd * p` cancels and one ulp of the clip coefficient is amplified),
state_dict exchange."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200.optim import FusedAdam
    DEV = torch.device("cuda:0")

SHAPES = [(2048, 4096), (2048,), (1, 1), (0,), (33, 7), (4097,), (256, 32)] + [(5, 3)] * 40      # > one argument table


def _params(seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [torch.randn(*s, device=DEV, generator=g).requires_grad_(True) for s in SHAPES]


@pytest.mark.parametrize("decoupled,wd,max_norm,scale", [(True, 1e-4, 1.0, 3.0), (True, 0.0, 1.0, 1e-4), (False, 1e-2, 1.0, 3.0),
                                                         (True, 1e-4, 0.0, 1.0)])
def test_fused_adam_matches_torch(decoupled, wd, max_norm, scale):
    pa, pb = _params(0), _params(0)
    ours = FusedAdam(pa, lr=1e-3, weight_decay=wd, decoupled=decoupled)
    ref = (torch.optim.AdamW if decoupled else torch.optim.Adam)(pb, lr=1e-3, weight_decay=wd)
    g = torch.Generator(device=DEV).manual_seed(1)
    for step in range(4):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=DEV, generator=g) * scale
            a.grad, b.grad = gr.clone(), gr.clone()
        if max_norm > 0:
            want_norm = torch.nn.utils.clip_grad_norm_(pb, max_norm)
            ref.step()
            norm = ours.clip_and_step(max_norm, want_norm=True)
            np.testing.assert_allclose(norm.item(), want_norm.item(), rtol=5e-6)
        else:
            ref.step()
            ours.step()
        for i, (a, b) in enumerate(zip(pa, pb)):
            np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-5, atol=0, err_msg=f"grad {i}")
            np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=1e-5, atol=5e-6, err_msg=f"p {i}")    # atol: 0.5 % of one lr-sized step (cancellation in g + wd p)
            if a.numel():
                sa, sb = ours.state[a], ref.state[b]
                np.testing.assert_allclose(sa["exp_avg"].cpu().numpy(), sb["exp_avg"].cpu().numpy(), rtol=1e-5, atol=1e-7)
                np.testing.assert_allclose(sa["exp_avg_sq"].cpu().numpy(), sb["exp_avg_sq"].cpu().numpy(), rtol=1e-5, atol=1e-9)
                assert float(sa["step"]) == float(sb["step"]) == step + 1
                assert a._version > step                       # in-place update is visible to version-keyed caches


def test_fused_adam_state_dict_exchange_and_scheduler():
    """A torch.optim.AdamW state loads into FusedAdam (and back) and training continues identically; LR schedulers drive
    param_groups['lr'] as usual."""
    pa, pb = _params(2)[:7], _params(2)[:7]
    ref = torch.optim.AdamW(pb, lr=1e-3, weight_decay=1e-4)
    g = torch.Generator(device=DEV).manual_seed(3)
    grads = [[torch.randn(a.shape, device=DEV, generator=g) for a in pa] for _ in range(4)]
    for step in range(2):
        for b, gr in zip(pb, grads[step]):
            b.grad = gr.clone()
        ref.step()
    ours = FusedAdam(pa, lr=5e-4, weight_decay=0.0)
    with torch.no_grad():
        for a, b in zip(pa, pb):
            a.copy_(b)
    import copy
    ours.load_state_dict(copy.deepcopy(ref.state_dict()))         # (torch hands out references to its own state tensors)
    assert ours.param_groups[0]["lr"] == 1e-3 and ours.param_groups[0]["weight_decay"] == 1e-4
    sched_a = torch.optim.lr_scheduler.LambdaLR(ours, lambda i: 0.5 ** i)
    sched_b = torch.optim.lr_scheduler.LambdaLR(ref, lambda i: 0.5 ** i)
    for step in range(2, 4):
        for a, b, gr in zip(pa, pb, grads[step]):
            a.grad, b.grad = gr.clone(), gr.clone()
        ours.step(); ref.step(); sched_a.step(); sched_b.step()
    for a, b in zip(pa, pb):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=1e-5, atol=1e-7)
    back = torch.optim.AdamW(pb, lr=1e-3)
    back.load_state_dict(copy.deepcopy(ours.state_dict()))           # and back
    assert float(back.state[pb[0]]["step"]) == 4.0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cpu = [torch.zeros(3, requires_grad=True)]
        cpu[0].grad = torch.ones(3)
        FusedAdam(cpu).step()


@pytest.mark.parametrize("kind", ["sgd", "adagrad", "rmsprop"])
@pytest.mark.parametrize("wd,max_norm", [(1e-4, 1.0), (0.0, 0.0)])
def test_fused_sgd_adagrad_rmsprop_match_torch(kind, wd, max_norm):
    """clip_grad_norm_ + torch.optim.SGD / Adagrad / RMSprop as index/trainer.py:62-75 builds them (lr + weight_decay only) against
    lcrec_simple_opt_clip_step over several steps; the optimiser state is interchangeable with torch's."""
    from lcrec_b200.optim import FusedSimple
    pa, pb = _params(3), _params(3)
    ours = FusedSimple(pa, kind, lr=1e-2, weight_decay=wd)
    ref = {"sgd": torch.optim.SGD, "adagrad": torch.optim.Adagrad, "rmsprop": torch.optim.RMSprop}[kind](pb, lr=1e-2, weight_decay=wd)
    g = torch.Generator(device=DEV).manual_seed(4)
    for step in range(4):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=DEV, generator=g) * 2.0
            a.grad, b.grad = gr.clone(), gr.clone()
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_(pb, max_norm)
        ref.step()
        ours.clip_and_step(max_norm)
        for a, b in zip(pa, pb):
            if a.numel():
                x, y = a.detach().cpu().numpy(), b.detach().cpu().numpy()
                bad = np.abs(x - y) > 1e-6 + 1e-5 * np.abs(y)
                # Adagrad / RMSprop divide by sqrt(accumulated g^2): where the clipped gradient and wd * p cancel (g_eff ~ 0) the
                # update is lr * sign-ish(g_eff) and one ulp of the clip coefficient decides it - a handful of elements in 8 M
                # (14 on a B200), each off by less than the step size; everything else agrees to 1e-5
                assert bad.mean() <= (1e-5 if kind != "sgd" else 0.0), (kind, int(bad.sum()), x.size)
                assert np.abs(x - y).max() <= (2e-2 * (step + 1) if kind != "sgd" else 1e-6)
                if max_norm > 0:
                    np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-5, atol=1e-7)
    key = {"sgd": None, "adagrad": "sum", "rmsprop": "square_avg"}[kind]
    if key:
        for a, b in zip(pa, pb):
            if a.numel():
                np.testing.assert_allclose(ours.state[a][key].cpu().numpy(), ref.state[b][key].cpu().numpy(), rtol=2e-5, atol=1e-7)
        sd = ours.state_dict()
        ref2 = type(ref)(_params(3), lr=1e-2, weight_decay=wd)
        ref2.load_state_dict(sd)                                    # torch's optimiser accepts the fused one's state
