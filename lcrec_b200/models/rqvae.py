"""RQ-VAE.  Import surface of reference ``index/models/rqvae.py`` (:10-85): same constructor,
attributes (``encoder``, ``rq``, ``decoder``), ``forward``, ``get_indices``, ``compute_loss`` and
state_dict keys, so reference checkpoints load unchanged."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from .layers import MLPLayers
from .rq import ResidualVectorQuantizer


class _ReconLossFn(torch.autograd.Function):
    """``F.mse_loss`` / ``F.l1_loss`` (mean) of rqvae.py:77-80 and its backward on lcrec_recon_loss / _backward."""

    @staticmethod
    def forward(ctx, out, xs, loss_type: str):
        from .. import ops
        ctx.save_for_backward(out, xs)
        ctx.loss_type = loss_type
        return ops.recon_loss(out.detach(), xs.detach(), loss_type)

    @staticmethod
    def backward(ctx, g):
        from .. import ops
        out, xs = ctx.saved_tensors
        go = ops.recon_loss_backward(out, xs, ctx.loss_type, g) if ctx.needs_input_grad[0] else None
        gx = None
        if ctx.needs_input_grad[1]:
            gx = -(go if go is not None else ops.recon_loss_backward(out, xs, ctx.loss_type, g))
        return go, gx, None


class RQVAE(nn.Module):
    _LOSSES = {"mse": F.mse_loss, "l1": F.l1_loss}
    _RQ = ResidualVectorQuantizer

    def __init__(self, in_dim=768, num_emb_list=None, e_dim=64, layers=None, dropout_prob=0.0, bn=False,
                 loss_type="mse", quant_loss_weight=1.0, beta=0.25, kmeans_init=False, kmeans_iters=100,
                 sk_epsilons=None, sk_iters=100):
        super().__init__()
        # every constructor argument is kept as an attribute of the same name (generate_indices.py and the checkpoint's
        # pickled args rely on them)
        for name, value in dict(in_dim=in_dim, num_emb_list=num_emb_list, e_dim=e_dim, layers=layers,
                                dropout_prob=dropout_prob, bn=bn, loss_type=loss_type,
                                quant_loss_weight=quant_loss_weight, beta=beta, kmeans_init=kmeans_init,
                                kmeans_iters=kmeans_iters, sk_epsilons=sk_epsilons, sk_iters=sk_iters).items():
            setattr(self, name, value)
        widths = [in_dim, *layers, e_dim]
        self.encode_layer_dims, self.decode_layer_dims = widths, widths[::-1]
        self.encoder = self._stack(widths)
        self.rq = self._quantizer()
        self.decoder = self._stack(widths[::-1])

    def _stack(self, widths):
        return MLPLayers(layers=widths, dropout=self.dropout_prob, bn=self.bn)

    def _quantizer(self, **extra):
        return self._RQ(self.num_emb_list, self.e_dim, beta=self.beta, kmeans_init=self.kmeans_init,
                                       kmeans_iters=self.kmeans_iters, sk_epsilons=self.sk_epsilons, sk_iters=self.sk_iters,
                                       **extra)

    def forward(self, x, use_sk=True):
        z = self.encoder(x)
        x_q, rq_loss, indices = self.rq(z, use_sk=use_sk)
        return self.decoder(x_q), rq_loss, indices

    @torch.no_grad()
    def get_indices(self, xs, use_sk=False):
        return self.rq(self.encoder(xs), use_sk=use_sk)[2]

    def compute_loss(self, out, quant_loss, xs=None):
        recon_fn = self._LOSSES.get(self.loss_type)
        if recon_fn is None:
            raise ValueError("incompatible loss type")           # rqvae.py:81
        if out.is_cuda and out.dtype == torch.float32 and xs.dtype == torch.float32 and out.shape == xs.shape:
            loss_recon = _ReconLossFn.apply(out, xs, self.loss_type)
        else:
            loss_recon = recon_fn(out, xs, reduction="mean")
        return loss_recon + self.quant_loss_weight * quant_loss, loss_recon
