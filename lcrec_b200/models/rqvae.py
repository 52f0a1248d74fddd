"""RQ-VAE.  Import surface of reference ``index/models/rqvae.py`` (:10-85): same constructor,
attributes (``encoder``, ``rq``, ``decoder``), ``forward``, ``get_indices``, ``compute_loss`` and
state_dict keys, so reference checkpoints load unchanged."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from .layers import MLPLayers
from .rq import ResidualVectorQuantizer


class RQVAE(nn.Module):
    def __init__(self, in_dim=768, num_emb_list=None, e_dim=64, layers=None, dropout_prob=0.0, bn=False,
                 loss_type="mse", quant_loss_weight=1.0, beta=0.25, kmeans_init=False, kmeans_iters=100,
                 sk_epsilons=None, sk_iters=100):
        super().__init__()
        self.in_dim = in_dim
        self.num_emb_list = num_emb_list
        self.e_dim = e_dim
        self.layers = layers
        self.dropout_prob = dropout_prob
        self.bn = bn
        self.loss_type = loss_type
        self.quant_loss_weight = quant_loss_weight
        self.beta = beta
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.sk_epsilons = sk_epsilons
        self.sk_iters = sk_iters

        self.encode_layer_dims = [in_dim] + list(layers) + [e_dim]
        self.encoder = MLPLayers(layers=self.encode_layer_dims, dropout=dropout_prob, bn=bn)
        self.rq = ResidualVectorQuantizer(num_emb_list, e_dim, beta=beta, kmeans_init=kmeans_init,
                                          kmeans_iters=kmeans_iters, sk_epsilons=sk_epsilons, sk_iters=sk_iters)
        self.decode_layer_dims = self.encode_layer_dims[::-1]
        self.decoder = MLPLayers(layers=self.decode_layer_dims, dropout=dropout_prob, bn=bn)

    def forward(self, x, use_sk=True):
        z = self.encoder(x)
        x_q, rq_loss, indices = self.rq(z, use_sk=use_sk)
        return self.decoder(x_q), rq_loss, indices

    @torch.no_grad()
    def get_indices(self, xs, use_sk=False):
        return self.rq(self.encoder(xs), use_sk=use_sk)[2]

    def compute_loss(self, out, quant_loss, xs=None):
        if self.loss_type == "mse":
            loss_recon = F.mse_loss(out, xs, reduction="mean")
        elif self.loss_type == "l1":
            loss_recon = F.l1_loss(out, xs, reduction="mean")
        else:
            raise ValueError("incompatible loss type")
        return loss_recon + self.quant_loss_weight * quant_loss, loss_recon
