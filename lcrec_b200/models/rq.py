"""Residual quantiser.  Import surface of reference ``index/models/rq.py`` (:7-56).

Without autograd (``get_indices``, evaluation) consecutive argmin levels are executed by ONE fused
kernel (``lcrec_rq_quantize``: codebooks staged in shared memory, residual kept in registers);
a level whose ``sk_epsilon`` is > 0 with ``use_sk`` goes through the Sinkhorn kernels.  With
autograd enabled the per-level modules are chained as in the reference so that gradients flow.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .vq import VectorQuantizer


# Training forward through ONE autograd node (lcrec_rq_train_forward / _backward) instead of the per-level torch glue;
# False restores the per-level modules (cross-checks).
FUSED_TRAIN = True


class _RQTrainFn(torch.autograd.Function):
    """(x_q, mean level loss, indices) of rq.py:39-56 with the analytic backward of the straight-through estimator:
    d x_q / d z = I, only level 0's commitment term reaches z, codebooks receive the codebook-loss term (summed in item
    order, deterministic).  Values and gradients equal the per-level autograd path to fp32 rounding (tested)."""

    @staticmethod
    def forward(ctx, z, rq, use_sk, *codebooks):
        d = rq.e_dim
        lat = z.detach().reshape(-1, d).float().contiguous()
        n = lat.shape[0]
        codes = rq.assign_codes(lat, use_sk)
        r = ops.rq_train_forward(lat, codes, [c.detach() for c in codebooks])
        mse = (r["sq_err"] / max(n * d, 1)).to(torch.float32)
        betas = [float(l.beta) for l in rq.vq_layers]
        if any(b != betas[0] for b in betas):                        # per-level betas: a (cached) device vector
            key = (tuple(betas), lat.device)
            if getattr(rq, "_beta_vec_key", None) != key:
                rq._beta_vec, rq._beta_vec_key = torch.tensor(betas, dtype=torch.float32, device=lat.device), key
            mean_loss = (mse + rq._beta_vec * mse).mean()
        else:
            mean_loss = (mse + betas[0] * mse).mean()                # codebook_loss + beta * commitment_loss, stack().mean()
        ctx.save_for_backward(r["diffs"], r["codes_t"])
        ctx.n_codes = [int(c.shape[0]) for c in codebooks]
        ctx.beta0 = float(rq.vq_layers[0].beta)
        ctx.z_shape = z.shape
        indices = codes.view(*z.shape[:-1], len(codebooks))
        ctx.mark_non_differentiable(indices)
        return r["xq"].view(z.shape), mean_loss, indices

    @staticmethod
    def backward(ctx, g_xq, g_loss, _g_idx):
        diffs, codes_t = ctx.saved_tensors
        need_z = ctx.needs_input_grad[0]
        need_cb = any(ctx.needs_input_grad[3:])
        gz, gcbs = ops.rq_train_backward(diffs, codes_t, ctx.n_codes, g_xq, g_loss, ctx.beta0, want_gz=need_z,
                                         want_gcb=need_cb)
        gz = gz.view(ctx.z_shape) if need_z else None
        return (gz, None, None) + (tuple(gcbs) if need_cb else (None,) * len(ctx.n_codes))


class ResidualVectorQuantizer(nn.Module):
    _LEVEL = VectorQuantizer

    def __init__(self, n_e_list, e_dim, sk_epsilons, beta=0.25, kmeans_init=False, kmeans_iters=100, sk_iters=100):
        super().__init__()
        self.n_e_list = n_e_list
        self.e_dim = e_dim
        self.num_quantizers = len(n_e_list)
        self.beta = beta
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.sk_epsilons = sk_epsilons
        self.sk_iters = sk_iters
        extra = self.__dict__.get("_level_extra", {})           # set by the EMA subclass before this constructor runs
        self.vq_layers = nn.ModuleList(
            self._LEVEL(n_e, e_dim, beta=beta, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                        sk_epsilon=eps, sk_iters=sk_iters, **extra)
            for n_e, eps in zip(n_e_list, sk_epsilons))

    def get_codebook(self):
        return torch.stack([q.get_codebook() for q in self.vq_layers])

    # ------------------------------------------------------------------ index selection for all levels, no autograd
    @torch.no_grad()
    def assign_codes(self, lat, use_sk=True):
        """(n, L) int64 codes of rq.py:45-52: runs of argmin levels in the fused kernel, Sinkhorn levels through
        VectorQuantizer.assign on the residual entering them."""
        layers = list(self.vq_layers)
        resid = lat
        codes = []
        i = 0
        while i < len(layers):
            j = i
            while j < len(layers) and not (use_sk and layers[j].sk_epsilon > 0):
                j += 1
            if j > i:
                last = j == len(layers)
                r = ops.rq_quantize(resid, [l.embedding.weight for l in layers[i:j]], resid_level=-1 if last else j - i)
                codes.append(r["codes"])
                resid = r["resid"]
                i = j
            if i < len(layers):
                l = layers[i]
                idx = l.assign(resid, use_sk=True)
                codes.append(idx[:, None])
                i += 1
                if i < len(layers):
                    q = l.embedding.weight.detach()[idx]
                    resid = resid - (resid + (q - resid))
        return torch.cat(codes, dim=1).contiguous()

    # ------------------------------------------------------------------ fused, no autograd
    @torch.no_grad()
    def quantize_fused(self, x, use_sk=True):
        """(x_q, mean_losses, indices) with the values of the reference forward, no graph."""
        layers = list(self.vq_layers)
        d = self.e_dim
        lat = x.reshape(-1, d).float().contiguous()
        n = lat.shape[0]
        resid = lat
        xq = torch.zeros_like(lat)
        codes, losses = [], []
        i = 0
        while i < len(layers):
            j = i
            while j < len(layers) and not (use_sk and layers[j].sk_epsilon > 0):
                j += 1
            if j > i:   # run of argmin levels [i, j)
                r = ops.rq_quantize(resid, [l.embedding.weight for l in layers[i:j]], resid_level=j - i,
                                    want_xq=True, want_sq_err=True)
                codes.append(r["codes"])
                xq = xq + r["xq"]
                resid = r["resid"]
                for l, e in zip(layers[i:j], r["sq_err"].tolist()):
                    mse = torch.tensor(e / max(n * d, 1), dtype=torch.float32, device=lat.device)
                    losses.append(mse + l.beta * mse)
                i = j
            if i < len(layers):   # one Sinkhorn level
                l = layers[i]
                idx = l.assign(resid, use_sk=True)
                q = l.embedding.weight.detach()[idx]
                mse = torch.mean((q - resid) ** 2)
                losses.append(mse + l.beta * mse)
                x_res = resid + (q - resid)
                resid = resid - x_res
                xq = xq + x_res
                codes.append(idx[:, None])
                i += 1
        indices = torch.cat(codes, dim=1).view(*x.shape[:-1], len(layers))
        return xq.view(x.shape), torch.stack(losses).mean(), indices

    def forward(self, x, use_sk=True):
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        pending_init = self.training and any(not q.initted for q in self.vq_layers)
        if not needs_graph and not pending_init:
            return self.quantize_fused(x, use_sk)
        if FUSED_TRAIN and not pending_init and type(self).forward is ResidualVectorQuantizer.forward:
            return _RQTrainFn.apply(x, self, use_sk, *[q.embedding.weight for q in self.vq_layers])
        return self._forward_levels(x, use_sk=use_sk)

    def _forward_levels(self, x, **level_kwargs):
        """The reference's chain of level modules (rq.py:45-55): autograd flows through every level."""
        losses, codes = [], []
        x_q, residual = 0, x
        for level in self.vq_layers:
            x_res, loss, idx = level(residual, **level_kwargs)
            residual, x_q = residual - x_res, x_q + x_res
            losses.append(loss)
            codes.append(idx)
        return x_q, torch.stack(losses).mean(), torch.stack(codes, dim=-1)
