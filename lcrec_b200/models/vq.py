"""One codebook level.  Import surface of reference ``index/models/vq.py`` (VectorQuantizer :7-99).

Index selection - the distance matrix, argmin, centring, fp64 Sinkhorn and argmax - runs in
liblcrec_b200.so.  What torch still does here is autograd plumbing for training: the embedding
gather (codebook gradient = scatter-add), the two MSE terms and the straight-through estimator.
``sk_epsilon`` may be mutated from outside (generate_indices.py:101-105 does) and is read at call time.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .layers import kmeans, sinkhorn_algorithm  # noqa: F401  (re-exported like the reference module)


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim, beta=0.25, kmeans_init=False, kmeans_iters=10, sk_epsilon=0.003, sk_iters=100):
        super().__init__()
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.sk_epsilon = sk_epsilon
        self.sk_iters = sk_iters
        self.embedding = nn.Embedding(n_e, e_dim)
        self.initted = not kmeans_init
        # data-parallel training: a lcrec_b200.distributed.DistributedSinkhorn makes the Sinkhorn step ONE problem over
        # the global batch (rows sharded over the ranks, column marginals all-reduced inside the kernel); None = local
        self.dist_sinkhorn = None
        if kmeans_init:
            self.embedding.weight.data.zero_()                       # filled by init_emb on the first batch
        else:
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    # ---- reference accessors (vq.py:29-49)
    def get_codebook(self):
        return self.embedding.weight

    def get_codebook_entry(self, indices, shape=None):
        z_q = self.embedding(indices)
        return z_q if shape is None else z_q.view(shape)

    def init_emb(self, data):
        self.embedding.weight.data.copy_(kmeans(data, self.n_e, self.kmeans_iters))
        self.initted = True

    @staticmethod
    def center_distance_for_constraint(distances):
        """vq.py:51-61 on the device; returns fp32 like the reference (the kernel itself emits the
        .double() the caller applies next, this wrapper rounds back for API fidelity - exact, the
        centred values are fp32 numbers)."""
        return ops.center_distances(distances).to(torch.float32)

    # ---- index selection on the device
    @torch.no_grad()
    def assign(self, latent, use_sk=True):
        """indices of vq.py:71-83 for ``latent`` (n, e_dim)."""
        cb = self.embedding.weight
        if not use_sk or self.sk_epsilon <= 0:
            return ops.rq_quantize(latent, [cb])["codes"][:, 0]
        d = ops.vq_distances(latent, cb)
        if self.dist_sinkhorn is not None:
            dc = self.dist_sinkhorn.center(d)                          # max / min over the global batch
            _, idx, flags = self.dist_sinkhorn(dc, self.sk_epsilon, self.sk_iters)
        else:
            dc = ops.center_distances(d)                              # fp64, raises AssertionError like vq.py:59
            idx, flags = ops.sinkhorn_dense_argmax(dc, self.sk_epsilon, self.sk_iters)     # cluster / DSMEM kernel
        ops.check_later("sinkhorn", flags)                            # peer-arrival error; NaN/Inf print of vq.py:81-82
        return idx

    def get_code(self, x, use_sk=True):
        """Named in the task statement; not a reference symbol.  Equals ``forward(x, use_sk)[2]``."""
        return self.assign(x.reshape(-1, self.e_dim), use_sk).view(x.shape[:-1])

    def forward(self, x, use_sk=True):
        latent = x.view(-1, self.e_dim)
        if not self.initted and self.training:
            self.init_emb(latent)
        indices = self.assign(latent.detach(), use_sk)
        x_q = self.embedding(indices).view(x.shape)
        commitment_loss = F.mse_loss(x_q.detach(), x)
        codebook_loss = F.mse_loss(x_q, x.detach())
        loss = codebook_loss + self.beta * commitment_loss
        x_q = x + (x_q - x).detach()                                  # straight-through, forward value x + (q - x)
        return x_q, loss, indices.view(x.shape[:-1])
