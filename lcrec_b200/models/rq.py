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


class ResidualVectorQuantizer(nn.Module):
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
        self.vq_layers = nn.ModuleList(
            VectorQuantizer(n_e, e_dim, beta=beta, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                            sk_epsilon=eps, sk_iters=sk_iters)
            for n_e, eps in zip(n_e_list, sk_epsilons))

    def get_codebook(self):
        return torch.stack([q.get_codebook() for q in self.vq_layers])

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
        all_losses, all_indices = [], []
        x_q = 0
        residual = x
        for quantizer in self.vq_layers:
            x_res, loss, indices = quantizer(residual, use_sk=use_sk)
            residual = residual - x_res
            x_q = x_q + x_res
            all_losses.append(loss)
            all_indices.append(indices)
        return x_q, torch.stack(all_losses).mean(), torch.stack(all_indices, dim=-1)
