"""RQ-VAE with EMA codebooks.  Import surface of reference ``index_improve/models/rqvae.py`` (:9-95)."""
from __future__ import annotations

import torch

from ...models.rqvae import RQVAE as _BaseRQVAE
from .rq import ResidualVectorQuantizer


class RQVAE(_BaseRQVAE):
    _RQ = ResidualVectorQuantizer

    def __init__(self, in_dim=768, num_emb_list=None, e_dim=64, layers=None, dropout_prob=0.0, bn=False,
                 loss_type="mse", quant_loss_weight=1.0, beta=0.25, kmeans_init=False, kmeans_iters=100,
                 sk_epsilons=None, sk_iters=100, ema_decay=0.99, epsilon=1e-5, reset_threshold=1e-5,
                 reset_interval=1000):
        self._ema_cfg = dict(ema_decay=ema_decay, epsilon=epsilon, reset_threshold=reset_threshold,
                             reset_interval=reset_interval)
        super().__init__(in_dim=in_dim, num_emb_list=num_emb_list, e_dim=e_dim, layers=layers,
                         dropout_prob=dropout_prob, bn=bn, loss_type=loss_type, quant_loss_weight=quant_loss_weight,
                         beta=beta, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters, sk_epsilons=sk_epsilons,
                         sk_iters=sk_iters)

    def _quantizer(self, **extra):
        return super()._quantizer(**self._ema_cfg, **extra)      # the EMA-aware ResidualVectorQuantizer of this package

    def forward(self, x, use_sk=True, use_ema=True):
        z = self.encoder(x)
        x_q, rq_loss, indices = self.rq(z, use_sk=use_sk, use_ema=use_ema)
        return self.decoder(x_q), rq_loss, indices

    @torch.no_grad()
    def get_indices(self, xs, use_sk=False):
        return self.rq(self.encoder(xs), use_sk=use_sk, use_ema=False)[2]

    def get_codebook_usage(self):
        return self.rq.get_codebook_usage()
