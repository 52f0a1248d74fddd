"""Residual quantiser over the EMA levels.  Import surface of reference ``index_improve/models/rq.py`` (:6-75)."""
from __future__ import annotations

import torch

from ...models.rq import ResidualVectorQuantizer as _BaseResidualVectorQuantizer
from .vq import VectorQuantizer


class ResidualVectorQuantizer(_BaseResidualVectorQuantizer):
    _LEVEL = VectorQuantizer

    def __init__(self, n_e_list, e_dim, sk_epsilons, beta=0.25, kmeans_init=False, kmeans_iters=100, sk_iters=100,
                 ema_decay=0.99, epsilon=1e-5, reset_threshold=1e-5, reset_interval=1000):
        object.__setattr__(self, "_level_extra", dict(ema_decay=ema_decay, epsilon=epsilon,
                                                      reset_threshold=reset_threshold, reset_interval=reset_interval))
        super().__init__(n_e_list, e_dim, sk_epsilons, beta=beta, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                         sk_iters=sk_iters)

    def forward(self, x, use_sk=True, use_ema=True):
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        pending_init = self.training and any(not q.initted for q in self.vq_layers)
        if not needs_graph and not pending_init and not (self.training and use_ema):
            return self.quantize_fused(x, use_sk)            # nothing to update: all levels in the fused kernels
        return self._forward_levels(x, use_sk=use_sk, use_ema=use_ema)

    def get_codebook_usage(self):
        usage_stats = []
        for i, quantizer in enumerate(self.vq_layers):
            stats = quantizer.get_codebook_usage()
            stats["quantizer_id"] = i
            usage_stats.append(stats)
        return usage_stats
