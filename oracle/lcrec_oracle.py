"""CPU oracle for the LC-Rec item-indexing hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference algorithm.  It is the checker
the CUDA path is compared with; nothing under ``lcrec_b200/`` may import it.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified
reference code itself, imported from /root/reference in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/``
(``tests/test_oracle_golden.py`` is the check).

Every function cites the reference lines it restates (paths relative to
/root/reference).  fp32 everywhere the reference is fp32, fp64 for Sinkhorn,
int64 indices, lowest index on ties.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
F64 = np.float64

TOKEN_PREFIX = ("<a_{}>", "<b_{}>", "<c_{}>", "<d_{}>", "<e_{}>")  # index/generate_indices.py:83


# --------------------------------------------------------------------------- #
# parameters
# --------------------------------------------------------------------------- #
@dataclass
class MlpParams:
    """Weights of one ``MLPLayers`` stack (index/models/layers.py:7-43).

    ``weights[i]`` is (out_i, in_i) like ``nn.Linear.weight``; ``bn`` holds, per
    non-final layer, (gamma, beta, running_mean, running_var, eps) or None.
    """
    weights: List[np.ndarray]
    biases: List[np.ndarray]
    bn: Optional[List[Optional[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, float]]]] = None


@dataclass
class RqvaeParams:
    encoder: MlpParams
    codebooks: List[np.ndarray]            # L arrays (K_l, D) fp32, rq.vq_layers.{l}.embedding.weight
    sk_epsilons: List[float]
    sk_iters: int = 50
    beta: float = 0.25
    decoder: Optional[MlpParams] = None
    quant_loss_weight: float = 1.0
    loss_type: str = "mse"


def params_from_state_dict(sd: Dict[str, np.ndarray], sk_epsilons: Sequence[float], sk_iters: int,
                           beta: float = 0.25, bn_eps: float = 1e-5,
                           quant_loss_weight: float = 1.0, loss_type: str = "mse") -> RqvaeParams:
    """Build oracle parameters from a reference ``RQVAE.state_dict()`` (numpy values).

    Key layout: ``encoder.mlp_layers.{i}.weight|bias`` for Linear,
    ``...{i}.running_mean`` etc. for BatchNorm1d, ``rq.vq_layers.{l}.embedding.weight``.
    """
    def mlp(prefix: str) -> Optional[MlpParams]:
        idxs = sorted({int(k.split(".")[2]) for k in sd if k.startswith(prefix + ".mlp_layers.")})
        if not idxs:
            return None
        lin, bns = [], {}
        for i in idxs:
            base = f"{prefix}.mlp_layers.{i}"
            if base + ".running_mean" in sd:
                bns[len(lin) - 1] = (sd[base + ".weight"].astype(F32), sd[base + ".bias"].astype(F32),
                                     sd[base + ".running_mean"].astype(F32),
                                     sd[base + ".running_var"].astype(F32), bn_eps)
            else:
                lin.append((sd[base + ".weight"].astype(F32), sd[base + ".bias"].astype(F32)))
        bn = [bns.get(j) for j in range(len(lin))] if bns else None
        return MlpParams([w for w, _ in lin], [b for _, b in lin], bn)

    n_levels = len({k.split(".")[2] for k in sd if k.startswith("rq.vq_layers.")})
    cbs = [sd[f"rq.vq_layers.{l}.embedding.weight"].astype(F32) for l in range(n_levels)]
    return RqvaeParams(encoder=mlp("encoder"), codebooks=cbs, sk_epsilons=list(sk_epsilons),
                       sk_iters=sk_iters, beta=beta, decoder=mlp("decoder"),
                       quant_loss_weight=quant_loss_weight, loss_type=loss_type)


# --------------------------------------------------------------------------- #
# a2: MLP
# --------------------------------------------------------------------------- #
def mlp_forward(x: np.ndarray, p: MlpParams, training_bn: bool = False) -> np.ndarray:
    """Eval-mode ``MLPLayers.forward`` (index/models/layers.py:18-43), relu activation.

    Per layer: Dropout (identity in eval / p=0) -> Linear -> [BatchNorm1d unless last]
    -> [ReLU unless last] (layers.py:22-30).  fp32 throughout like ``nn.Linear``.
    """
    h = np.ascontiguousarray(x, dtype=F32)
    last = len(p.weights) - 1
    for i, (w, b) in enumerate(zip(p.weights, p.biases)):
        h = h @ w.T.astype(F32) + b.astype(F32)
        if i != last:
            if p.bn is not None and p.bn[i] is not None:
                g, bt, mu, var, eps = p.bn[i]
                if training_bn:  # batch statistics, biased variance (torch BatchNorm1d training)
                    mu = h.mean(axis=0, dtype=F64).astype(F32)
                    var = h.var(axis=0, dtype=F64).astype(F32)
                h = (h - mu) / np.sqrt(var + F32(eps)).astype(F32) * g + bt
                h = h.astype(F32)
            h = np.maximum(h, F32(0))
    return h.astype(F32)


# --------------------------------------------------------------------------- #
# a4/a5/a6/a7: one quantiser level
# --------------------------------------------------------------------------- #
def fma32(a, b, c):
    """Correctly rounded fp32 ``fma(a, b, c)`` on arrays: the product of two fp32 numbers is exact in fp64; the fp64 sum
    ``t`` and its exact error ``e`` (TwoSum) decide the single rounding to fp32, including the double-rounding corner
    (``t`` exactly half-way between two fp32 numbers while ``e != 0``)."""
    a64, b64, c64 = (np.asarray(v, F32).astype(F64) for v in (a, b, c))
    p = a64 * b64
    t = p + c64
    r = t.astype(F32)
    # t is an exact fp32 midpoint iff its low 29 mantissa bits are 1000...0; only then can the second rounding go wrong
    mid = (np.ascontiguousarray(t).view(np.int64) & 0x1FFFFFFF) == 0x10000000
    if not mid.any():
        return r
    bb = t - p
    e = (p - (t - bb)) + (c64 - bb)
    r64 = r.astype(F64)
    with np.errstate(invalid="ignore", over="ignore"):
        up = np.nextafter(r, F32(np.inf)).astype(F64)
        dn = np.nextafter(r, F32(-np.inf)).astype(F64)
        tie_up = mid & (t > r64) & ((t - r64) == (up - t)) & (e > 0)      # true value beyond the midpoint towards `up`
        tie_dn = mid & (t < r64) & ((r64 - t) == (t - dn)) & (e < 0)
    return np.where(tie_up, up.astype(F32), np.where(tie_dn, dn.astype(F32), r))


def _chain_dot(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``out[i, k] = fma(a[i, D-1], b[k, D-1], ... fma(a[i, 0], b[k, 0], 0))``: fp32 sequential chain, d ascending."""
    acc = np.zeros((a.shape[0], b.shape[0]), dtype=F32)
    for d in range(a.shape[1]):
        acc = fma32(a[:, d:d + 1], b[None, :, d], acc)
    return acc


def _chain_sqnorm(a: np.ndarray) -> np.ndarray:
    acc = np.zeros(a.shape[0], dtype=F32)
    for d in range(a.shape[1]):
        acc = fma32(a[:, d], a[:, d], acc)
    return acc


def vq_distances(latent: np.ndarray, codebook: np.ndarray, order: str = "blas") -> np.ndarray:
    """``d = sum(x^2) + sum(c^2)^T - 2 x c^T`` in fp32, evaluation order (xx + cc) - 2 dot
    (index/models/vq.py:71-73).

    The reference leaves the fp32 summation ORDER inside ``torch.sum`` / ``torch.matmul`` to the library (MKL on CPU,
    cuBLAS on CUDA; it changes with the batch shape and the thread count), so distances are defined up to a few ulp.
    ``order="blas"``: numpy pairwise sums + OpenBLAS sgemm (one such library order).
    ``order="chain"``: every sum as one fp32 fma chain in ascending dimension - the order the CUDA kernels use
    (``csrc/rq_fused.cu``, ``csrc/sinkhorn.cu``); with it the oracle is comparable bit for bit with the device path."""
    latent = latent.astype(F32, copy=False)
    codebook = codebook.astype(F32, copy=False)
    if order == "chain":
        xx = _chain_sqnorm(latent)[:, None]
        cc = _chain_sqnorm(codebook)[None, :]
        dot = _chain_dot(latent, codebook)
        return ((xx + cc).astype(F32) - F32(2) * dot).astype(F32)
    xx = np.sum(latent * latent, axis=1, keepdims=True, dtype=F32)
    cc = np.sum(codebook * codebook, axis=1, keepdims=True, dtype=F32).T
    dot = latent @ codebook.T
    return ((xx + cc) - F32(2) * dot).astype(F32)


def center_distance_for_constraint(d: np.ndarray) -> np.ndarray:
    """Global max/min centring over the whole (B, K) matrix, fp32 (index/models/vq.py:51-61)."""
    d = d.astype(F32, copy=False)
    mx = d.max()
    mn = d.min()
    middle = F32((mx + mn) / F32(2))
    amplitude = F32(F32(mx - middle) + F32(1e-5))
    if not amplitude > 0:
        raise AssertionError("amplitude > 0")  # vq.py:59
    return ((d - middle) / amplitude).astype(F32)


def sinkhorn_algorithm(distances: np.ndarray, epsilon: float, sinkhorn_iterations: int) -> np.ndarray:
    """Literal restatement of ``sinkhorn_algorithm`` (index/models/layers.py:85-108).

    The code (not its swapped comments) is authoritative: per iteration divide by the
    per-row sum (dim=1), by B, by the per-column sum (dim=0), by K; finally multiply by B.
    Works in the dtype of ``distances`` (fp64 at the call site vq.py:78-79).
    """
    dist = np.asarray(distances)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore", under="ignore"):
        q = np.exp(-dist / epsilon)
        n_rows, n_cols = q.shape
        q = q / q.sum(axis=-1, keepdims=True).sum(axis=-2, keepdims=True)
        for _ in range(sinkhorn_iterations):
            q = q / q.sum(axis=1, keepdims=True)
            q = q / n_rows
            q = q / q.sum(axis=0, keepdims=True)
            q = q / n_cols
        q = q * n_rows
    return q


def _argmax_first_nan_wins(q: np.ndarray) -> np.ndarray:
    """``torch.argmax`` semantics: first maximal index; a NaN counts as the maximum."""
    nan = np.isnan(q)
    idx = np.argmax(np.where(nan, np.inf, q), axis=-1)
    has_nan = nan.any(axis=-1)
    if has_nan.any():
        idx = np.where(has_nan, np.argmax(nan, axis=-1), idx)
    return idx.astype(np.int64)


def vq_assign(latent: np.ndarray, codebook: np.ndarray, use_sk: bool, sk_epsilon: float,
              sk_iters: int, want_q: bool = False, order: str = "blas"):
    """Index selection of ``VectorQuantizer.forward`` (index/models/vq.py:71-83)."""
    d = vq_distances(latent, codebook, order)
    if (not use_sk) or sk_epsilon <= 0:
        idx = np.argmin(d, axis=-1).astype(np.int64)   # first minimum, like torch.argmin
        return (idx, d, None) if want_q else idx
    dc = center_distance_for_constraint(d).astype(F64)
    q = sinkhorn_algorithm(dc, sk_epsilon, sk_iters)
    idx = _argmax_first_nan_wins(q)
    return (idx, d, q) if want_q else idx


def _mse(a: np.ndarray, b: np.ndarray) -> np.float32:
    """``F.mse_loss(..., reduction='mean')`` in fp32 (accumulated in fp64 here, rounded once)."""
    diff = a.astype(F32) - b.astype(F32)
    return F32(np.mean((diff * diff).astype(F32), dtype=F64))


def vq_forward(x: np.ndarray, codebook: np.ndarray, use_sk: bool, sk_epsilon: float, sk_iters: int,
               beta: float, order: str = "blas"):
    """Forward values of ``VectorQuantizer.forward`` (index/models/vq.py:63-99).

    Returns (x_q straight-through forward value ``x + (q - x)``, loss, indices).
    """
    x = x.astype(F32, copy=False)
    idx = vq_assign(x, codebook, use_sk, sk_epsilon, sk_iters, order=order)
    q = codebook.astype(F32)[idx]
    mse = _mse(q, x)
    loss = F32(mse + F32(beta) * mse)          # codebook_loss + beta * commitment_loss (vq.py:90-92)
    x_q = (x + (q - x)).astype(F32)            # vq.py:95
    return x_q, loss, idx


# --------------------------------------------------------------------------- #
# a11: clip_grad_norm_ + Adam / AdamW step (index/trainer.py:49-81, :117-118 -> torch.nn.utils / torch.optim)
# --------------------------------------------------------------------------- #
def adam_clip_step(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                   weight_decay: float = 0.0, decoupled: bool = True, max_norm: float = 0.0):
    """One ``clip_grad_norm_(params, max_norm)`` + ``optimizer.step()`` on lists of fp32 arrays (updated copies returned):
    total norm over all tensors, coef = min(1, max_norm / (norm + 1e-6)), g *= coef; AdamW: p *= 1 - lr wd / Adam: g += wd p;
    m += (g - m)(1 - b1); v = v b2 + (1 - b2) g g; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
    ``step`` is the count AFTER this update.  Returns (params, clipped grads, exp_avg, exp_avg_sq, total_norm)."""
    b1, b2 = betas
    g = [x.astype(F32, copy=True) for x in grads]
    norm = F32(np.sqrt(sum(float(np.sum(x.astype(F64) ** 2)) for x in g)))
    if max_norm > 0:
        coef = F32(max_norm) / F32(norm + F32(1e-6))
        coef = coef if coef < 1 else F32(1.0)
        g = [(x * coef).astype(F32) for x in g]
    step_size = F32(lr / (1.0 - b1 ** step))
    bc2_sqrt = F32(np.sqrt(1.0 - b2 ** step))
    out_p, out_m, out_v = [], [], []
    for p, gi, m, v in zip(params, g, exp_avg, exp_avg_sq):
        p = p.astype(F32, copy=True)
        if decoupled:
            p = (p * F32(1.0 - lr * weight_decay)).astype(F32)
            gu = gi
        else:
            gu = (gi + F32(weight_decay) * p).astype(F32) if weight_decay != 0 else gi
        m = (m + (gu - m) * F32(1.0 - b1)).astype(F32)
        v = (v * F32(b2) + (gu * gu) * F32(1.0 - b2)).astype(F32)
        denom = (np.sqrt(v) / bc2_sqrt + F32(eps)).astype(F32)
        p = (p - step_size * (m / denom)).astype(F32)
        out_p.append(p); out_m.append(m); out_v.append(v)
    return out_p, g, out_m, out_v, norm


# --------------------------------------------------------------------------- #
# f1: k-means++ seeding of scikit-learn with PRE-DRAWN random numbers (groundwork for a device seeding kernel)
# --------------------------------------------------------------------------- #
def kmeanspp_draws(random_state, n_clusters: int):
    """The random numbers ``sklearn.cluster.kmeans_plusplus`` (1.9.0, `_kmeans_plusplus`) consumes, in its order: one
    ``random_sample()`` for the first centre (``RandomState.choice(n, p=...)`` draws exactly one uniform), then
    ``uniform(size=2 + int(log(K)))`` per further centre.  Their number does not depend on the data, so a device seeding
    kernel can receive them up front and still consume numpy's global RNG exactly like the reference's KMeans.fit."""
    trials = 2 + int(np.log(n_clusters))
    u0 = float(random_state.random_sample())
    return u0, [random_state.uniform(size=trials) for _ in range(1, n_clusters)]


def kmeanspp_predrawn(xc: np.ndarray, n_clusters: int, u0: float, draws) -> np.ndarray:
    """Indices of the seeds `_kmeans_plusplus` picks on the centred fp32 batch ``xc`` with unit sample weights, given the
    pre-drawn numbers: first centre = searchsorted(cumsum(1/n) / last, u0, 'right') (``RandomState.choice``); then per centre:
    candidates = searchsorted(fp32 sequential cumsum of the closest squared distances, u * potential), squared distances of
    the candidates to all rows in fp64 from the up-cast rows (``-2 x.y + |x|^2 + |y|^2``, clipped at 0, rounded to fp32:
    sklearn's `_euclidean_distances_upcast`), the candidate with the smallest new potential wins (first on ties)."""
    x32 = np.ascontiguousarray(xc, dtype=F32)
    n = x32.shape[0]
    x64 = x32.astype(F64)
    norms = (x64 * x64).sum(axis=1)
    ones = np.ones(n, dtype=F32)

    def sqdist(rows):                                   # (len(rows), n) fp32
        d = -2.0 * (x64[rows] @ x64.T)
        d += norms[rows][:, None]
        d += norms[None, :]
        return np.maximum(d.astype(F32), F32(0))

    p = ones.astype(F64) / F64(ones.sum())
    cdf = p.cumsum()
    cdf /= cdf[-1]
    first = int(cdf.searchsorted(u0, side="right"))
    indices = [first]
    closest = sqdist([first])[0]
    pot = closest @ ones
    for u in draws:
        rand_vals = u * pot
        cand = np.searchsorted(np.cumsum(ones * closest), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        dc = sqdist(cand)
        np.minimum(closest, dc, out=dc)
        cand_pot = dc @ ones.reshape(-1, 1)
        best = int(np.argmin(cand_pot))
        pot = cand_pot[best]
        closest = dc[best]
        indices.append(int(cand[best]))
    return np.asarray(indices, dtype=np.int64)


# --------------------------------------------------------------------------- #
# f3: index_improve EMA codebook update, usage statistics, dead-code reset
# --------------------------------------------------------------------------- #
def _fma32(a, b, c):
    """fp32 ``fma(a, b, c)``: the product of two fp32 numbers is exact in fp64, so one fp64 add + one rounding to
    fp32 differs from the fused result only by double rounding (never seen against the golden vectors).  torch's CPU
    ``add_(t, alpha=s)`` is ``vec::fmadd(t, s, self)`` and the CUDA functor ``a + alpha * b`` contracts to FFMA."""
    return (np.asarray(a, F32).astype(F64) * np.asarray(b, F32).astype(F64) + np.asarray(c, F32).astype(F64)).astype(F32)


@dataclass
class EmaState:
    """Buffers of the index_improve ``VectorQuantizer`` (index_improve/models/vq.py:39-42)."""
    codebook: np.ndarray          # (K, D) fp32  embedding.weight
    cluster_size: np.ndarray      # (K,)  fp32   _ema_cluster_size
    ema_w: np.ndarray             # (K, D) fp32  _ema_w
    step_count: int = 0


def ema_update(latent: np.ndarray, indices: np.ndarray, st: EmaState, ema_decay: float = 0.99,
               epsilon: float = 1e-5) -> EmaState:
    """The ``self.training and use_ema`` block of the improved quantiser (index_improve/models/vq.py:146-187).

    Scalars are Python doubles rounded to fp32 when they meet an fp32 tensor; ``index_add_`` on the CPU adds the
    rows of a code in ascending item order (``np.add.at`` does the same); the convex update of the codebook is three
    separate roundings (mul, mul, add), only for codes whose smoothed count exceeds ``epsilon``.
    """
    latent = latent.astype(F32, copy=False).reshape(-1, st.codebook.shape[1])
    idx = np.asarray(indices).reshape(-1).astype(np.int64)
    k, d = st.codebook.shape
    decay32, alpha32 = F32(ema_decay), F32(1 - ema_decay)
    counts = np.bincount(idx, minlength=k).astype(F32)                                   # vq.py:151-152
    cs = _fma32(counts, alpha32, (st.cluster_size.astype(F32) * decay32).astype(F32))    # vq.py:155-157
    dw = np.zeros((k, d), dtype=F32)
    np.add.at(dw, idx, latent)                                                           # vq.py:163-167, item order
    w = _fma32(dw, alpha32, (st.ema_w.astype(F32) * decay32).astype(F32))                # vq.py:169
    normalized = (w / (cs[:, None] + F32(epsilon)).astype(F32)).astype(F32)              # vq.py:173
    used = cs > F32(epsilon)                                                             # vq.py:176
    update_rate = 1 - ema_decay
    cb = st.codebook.astype(F32, copy=True)
    cb[used] = ((cb[used] * F32(1 - update_rate)).astype(F32)
                + (normalized[used] * F32(update_rate)).astype(F32)).astype(F32)         # vq.py:181-184
    return EmaState(cb, cs, w, st.step_count + 1)


def codebook_usage(cluster_size: np.ndarray, epsilon: float = 1e-5, reset_threshold: float = 1e-5):
    """``get_codebook_usage`` (index_improve/models/vq.py:205-217): (utilization, used_codes, total_codes)."""
    cs = cluster_size.astype(F32)
    total = F32(cs.sum(dtype=F32) + F32(epsilon))
    usage = (cs / total).astype(F32)
    used = int((usage > F32(reset_threshold)).sum())
    return used / cs.shape[0], used, int(cs.shape[0])


def unused_codes(cluster_size: np.ndarray, epsilon: float = 1e-5, reset_threshold: float = 1e-5) -> np.ndarray:
    """Indices ``_reset_unused_codes`` would consider dead (index_improve/models/vq.py:83-90); the replacement
    vectors themselves are drawn from torch's RNG (randint / randperm / randn_like) and are not restated."""
    cs = cluster_size.astype(F32)
    usage = (cs / F32(cs.sum(dtype=F32) + F32(epsilon))).astype(F32)
    return np.nonzero(usage < F32(reset_threshold))[0].astype(np.int64)


def vq_forward_ema(x: np.ndarray, st: EmaState, use_sk: bool, sk_epsilon: float, sk_iters: int, beta: float,
                   ema_decay: float = 0.99, epsilon: float = 1e-5, use_ema: bool = True, training: bool = True):
    """Forward values of the improved quantiser for one step that does not cross ``reset_interval``
    (index_improve/models/vq.py:115-203): the loss and x_q use the codebook BEFORE the EMA step."""
    x_q, loss, idx = vq_forward(x, st.codebook, use_sk, sk_epsilon, sk_iters, beta)
    if training and use_ema:
        st = ema_update(x, idx, st, ema_decay, epsilon)
    return x_q, loss, idx, st


# --------------------------------------------------------------------------- #
# f4: upstream embedding producer hand-off (masked mean pool of the PLM's last hidden state)
# --------------------------------------------------------------------------- #
def masked_mean_pool(hidden: np.ndarray, attention_mask: np.ndarray) -> np.ndarray:
    """``(h * mask[..., None]).sum(1) / mask.sum(-1, keepdim=True)`` (data_process/amazon_text_emb.py:91-92), fp32;
    rows are added in position order."""
    h = hidden.astype(F32, copy=False)
    m = np.asarray(attention_mask)
    acc = np.zeros((h.shape[0], h.shape[2]), dtype=F32)
    for t in range(h.shape[1]):
        acc = (acc + (h[:, t, :] * m[:, t, None].astype(F32)).astype(F32)).astype(F32)
    return (acc / m.sum(axis=-1, keepdims=True).astype(F32)).astype(F32)


def item_embedding(field_hiddens: Sequence[np.ndarray], field_masks: Sequence[np.ndarray]) -> np.ndarray:
    """Mean over the text fields of the per-field pools: ``torch.stack(field_embeddings, 0).mean(0)``
    (data_process/amazon_text_emb.py:96)."""
    pools = [masked_mean_pool(h, m) for h, m in zip(field_hiddens, field_masks)]
    acc = pools[0]
    for q in pools[1:]:
        acc = (acc + q).astype(F32)
    return (acc / F32(len(pools))).astype(F32)


# --------------------------------------------------------------------------- #
# a9/a10: residual quantiser and the model
# --------------------------------------------------------------------------- #
def rq_forward(z: np.ndarray, p: RqvaeParams, use_sk: bool, order: str = "blas"):
    """``ResidualVectorQuantizer.forward`` (index/models/rq.py:39-56).  ``order``: fp32 summation order of the distances
    (see ``vq_distances``)."""
    residual = z.astype(F32, copy=True)
    x_q = np.zeros_like(residual)
    losses, indices = [], []
    for cb, eps in zip(p.codebooks, p.sk_epsilons):
        x_res, loss, idx = vq_forward(residual, cb, use_sk, eps, p.sk_iters, p.beta, order)
        residual = (residual - x_res).astype(F32)
        x_q = (x_q + x_res).astype(F32)
        losses.append(loss)
        indices.append(idx)
    mean_loss = F32(np.mean(np.array(losses, dtype=F32), dtype=F64))
    return x_q, mean_loss, np.stack(indices, axis=-1)


def rq_trace(z: np.ndarray, p: RqvaeParams):
    """argmin-only pass that also returns the residual entering every level (teacher forcing)."""
    residual = z.astype(F32, copy=True)
    resids, dists, idxs = [], [], []
    for cb in p.codebooks:
        resids.append(residual.copy())
        d = vq_distances(residual, cb)
        idx = np.argmin(d, axis=-1).astype(np.int64)
        q = cb.astype(F32)[idx]
        x_res = (residual + (q - residual)).astype(F32)
        residual = (residual - x_res).astype(F32)
        dists.append(d)
        idxs.append(idx)
    return resids, dists, np.stack(idxs, axis=-1)


def get_indices(x: np.ndarray, p: RqvaeParams, use_sk: bool = False) -> np.ndarray:
    """``RQVAE.get_indices`` (index/models/rqvae.py:68-72)."""
    z = mlp_forward(x, p.encoder)
    return rq_forward(z, p, use_sk)[2]


def rqvae_forward(x: np.ndarray, p: RqvaeParams, use_sk: bool = True):
    """``RQVAE.forward`` + ``compute_loss`` forward values (index/models/rqvae.py:61-85)."""
    z = mlp_forward(x, p.encoder)
    x_q, rq_loss, idx = rq_forward(z, p, use_sk)
    out = mlp_forward(x_q, p.decoder)
    if p.loss_type == "mse":
        recon = _mse(out, x)
    elif p.loss_type == "l1":
        recon = F32(np.mean(np.abs(out.astype(F32) - x.astype(F32)), dtype=F64))
    else:
        raise ValueError("incompatible loss type")   # rqvae.py:81
    total = F32(recon + F32(p.quant_loss_weight) * rq_loss)
    return out, rq_loss, idx, total, recon


# --------------------------------------------------------------------------- #
# a12/a14: collision bookkeeping
# --------------------------------------------------------------------------- #
def collision_groups(codes: np.ndarray) -> List[List[int]]:
    """``get_collision_item`` (index/generate_indices.py:29-42): lists of item ids sharing an
    identical code tuple, groups in first-occurrence order, members ascending."""
    table: Dict[bytes, List[int]] = {}
    c = np.ascontiguousarray(codes, dtype=np.int64)
    for i in range(c.shape[0]):
        table.setdefault(c[i].tobytes(), []).append(i)
    return [g for g in table.values() if len(g) > 1]


def n_unique_codes(codes: np.ndarray) -> int:
    return int(np.unique(np.ascontiguousarray(codes, dtype=np.int64), axis=0).shape[0])


def collision_rate(codes: np.ndarray) -> float:
    """(N - |unique|) / N (index/trainer.py:150, index/generate_indices.py:133-136)."""
    n = codes.shape[0]
    return (n - n_unique_codes(codes)) / n


def max_conflicts(codes: np.ndarray) -> int:
    """``max(get_indices_count(...).values())`` (index/generate_indices.py:23-27,132)."""
    _, cnt = np.unique(np.ascontiguousarray(codes, dtype=np.int64), axis=0, return_counts=True)
    return int(cnt.max())


# --------------------------------------------------------------------------- #
# a15/a16: generate_indices
# --------------------------------------------------------------------------- #
@dataclass
class GenTrace:
    codes_pass0: np.ndarray
    rounds: List[np.ndarray] = field(default_factory=list)   # code table after each round
    n_groups: List[int] = field(default_factory=list)
    n_rows: List[int] = field(default_factory=list)


def generation_epsilons(p: RqvaeParams) -> List[float]:
    """index/generate_indices.py:101-105: levels 0..L-2 -> 0; last keeps its value, 0 -> 0.003."""
    eps = [0.0] * (len(p.codebooks) - 1)
    last = p.sk_epsilons[-1]
    eps.append(0.003 if last == 0.0 else last)
    return eps


def generate_indices(x: np.ndarray, p: RqvaeParams, batch_size: int = 64, max_rounds: int = 20,
                     reencode: bool = True) -> Tuple[np.ndarray, GenTrace]:
    """``index/generate_indices.py:85-128``.

    PASS 0: argmin codes for all items in batches of ``batch_size`` (:85-95).
    Then up to ``max_rounds`` rounds: every collision group is re-quantised on its own with
    Sinkhorn on the last level, all L codes of its members overwritten (:107-128).

    ``reencode=True`` follows the reference literally (encoder re-run on the group's rows);
    ``reencode=False`` re-uses the latents of PASS 0 (what the CUDA path does: same maths, the
    difference is GEMM batch-shape rounding only).
    """
    x = np.ascontiguousarray(x, dtype=F32)
    n = x.shape[0]
    codes = np.empty((n, len(p.codebooks)), dtype=np.int64)
    latents = np.empty((n, p.codebooks[0].shape[1]), dtype=F32) if not reencode else None
    for s in range(0, n, batch_size):
        z = mlp_forward(x[s:s + batch_size], p.encoder)
        if latents is not None:
            latents[s:s + batch_size] = z
        codes[s:s + batch_size] = rq_forward(z, p, use_sk=False)[2]
    trace = GenTrace(codes_pass0=codes.copy())

    p_sk = RqvaeParams(encoder=p.encoder, codebooks=p.codebooks, sk_epsilons=generation_epsilons(p),
                       sk_iters=p.sk_iters, beta=p.beta)
    tt = 0
    while tt < max_rounds and n_unique_codes(codes) != n:
        groups = collision_groups(codes)
        rows = 0
        for g in groups:
            z = mlp_forward(x[g], p.encoder) if reencode else latents[g]
            codes[g] = rq_forward(z, p_sk, use_sk=True)[2]
            rows += len(g)
        trace.rounds.append(codes.copy())
        trace.n_groups.append(len(groups))
        trace.n_rows.append(rows)
        tt += 1
    return codes, trace


def resolve_collisions(codes_pass0: np.ndarray, resid_last: np.ndarray, codebook_last: np.ndarray,
                       sk_epsilon: float, sk_iters: int, max_rounds: int = 20, order: str = "blas") -> Tuple[np.ndarray, GenTrace]:
    """The collision rounds of ``index/generate_indices.py:107-128`` started from GIVEN inputs of the last
    level: the PASS-0 table and the residual entering the last quantiser (``rq.py:45-48``).

    Every group is an independent Sinkhorn problem on its own rows (``vq.py:71-83``).  With the latents
    held fixed, levels 0..L-2 of a re-quantised row are its PASS-0 argmin codes again (their epsilon is
    forced to 0, ``generate_indices.py:101-103``), so only the last column changes.  This is the form the
    parity tests use to separate the LOOP (integer bookkeeping + fp64 Sinkhorn: must be bit-exact on
    equal inputs) from the ENCODER (fp32 GEMMs: equal within 1e-5 relative, never bitwise - the reference
    itself re-encodes every group at a different batch shape).
    """
    codes = np.array(codes_pass0, dtype=np.int64, copy=True)
    n = codes.shape[0]
    trace = GenTrace(codes_pass0=codes.copy())
    tt = 0
    while tt < max_rounds and n_unique_codes(codes) != n:
        groups = collision_groups(codes)
        rows = 0
        for g in groups:
            codes[g, -1] = vq_assign(resid_last[g], codebook_last, True, sk_epsilon, sk_iters, order=order)
            rows += len(g)
        trace.rounds.append(codes.copy())
        trace.n_groups.append(len(groups))
        trace.n_rows.append(rows)
        tt += 1
    return codes, trace


@dataclass
class LoopLedger:
    """Per-row account of how two runs of the collision loop differ."""
    rows_differing_final: int = 0
    first_round: int = -1                      # first round (1-based) after which any row differs
    primary: List[Tuple[int, int]] = field(default_factory=list)   # (round, item): same group + same incoming table, other pick
    cascade: int = 0                           # rows whose first difference comes from a group that already differed
    per_round: List[int] = field(default_factory=list)


def loop_ledger(tables_a: Sequence[np.ndarray], tables_b: Sequence[np.ndarray]) -> LoopLedger:
    """Compare two per-round table sequences ``[pass0, after round 1, ...]`` of the same items.

    A row's FIRST difference (round r) is *primary* when the group that re-quantised it in round r had
    the same members and the same incoming codes in both runs - the two sides then solved the same
    Sinkhorn problem on (nominally) the same rows and picked differently, which only input rounding or a
    kernel defect can cause; it is a *cascade* when the group itself already differed (a consequence of an
    earlier primary difference).  Shorter sequences are padded with their last table (a run that stopped).
    """
    ta, tb = list(tables_a), list(tables_b)
    m = max(len(ta), len(tb))
    ta += [ta[-1]] * (m - len(ta)); tb += [tb[-1]] * (m - len(tb))
    led = LoopLedger()
    n = ta[0].shape[0]
    seen = np.zeros(n, dtype=bool) | (ta[0] != tb[0]).any(axis=1)
    led.per_round.append(int(seen.sum()))
    for r in range(1, m):
        diff = (ta[r] != tb[r]).any(axis=1)
        led.per_round.append(int(diff.sum()))
        new = np.nonzero(diff & ~seen)[0]
        if len(new) and led.first_round < 0:
            led.first_round = r
        if len(new):
            ga = {g[0]: g for g in collision_groups(ta[r - 1])}
            gb = {g[0]: g for g in collision_groups(tb[r - 1])}
            owner_a = {}
            for g in ga.values():
                for i in g:
                    owner_a[i] = g
            for i in new:
                g = owner_a.get(int(i))
                same = g is not None and gb.get(g[0]) == g and bool((ta[r - 1][g] == tb[r - 1][g]).all())
                if same:
                    led.primary.append((r, int(i)))
                else:
                    led.cascade += 1
        seen |= diff
    led.rows_differing_final = int((ta[-1] != tb[-1]).any(axis=1).sum())
    return led


def distance_ulp_witness(resid_a: np.ndarray, resid_b: np.ndarray, codebook: np.ndarray, order_a: str = "blas",
                         order_b: str = "blas") -> Tuple[float, float]:
    """For one group solved on two roundings of the same rows: (max difference of the fp32 distance matrices
    ``vq.py:71-73`` in ulps of the entry, max relative difference of the residual rows).  The pair is the
    constructive witness "the other pick follows from an N-ulp perturbation of the distances"."""
    da, db = vq_distances(resid_a, codebook, order_a), vq_distances(resid_b, codebook, order_b)
    ulp = np.spacing(np.maximum(np.abs(da), np.abs(db)).astype(F32)).astype(F64)
    n_ulp = float(np.max(np.abs(da.astype(F64) - db.astype(F64)) / ulp))
    scale = max(float(np.abs(resid_a).max()), 1e-30)
    return n_ulp, float(np.abs(resid_a.astype(F64) - resid_b.astype(F64)).max() / scale)


def codes_to_tokens(codes: np.ndarray) -> List[List[str]]:
    """``prefix[i].format(int(ind))`` (index/generate_indices.py:83,90-92)."""
    return [[TOKEN_PREFIX[i].format(int(v)) for i, v in enumerate(row)] for row in codes]


def index_json(codes: np.ndarray) -> str:
    """Exactly what ``json.dump({item: [tokens]})`` writes (index/generate_indices.py:138-145)."""
    return json.dumps({i: toks for i, toks in enumerate(codes_to_tokens(codes))})


# --------------------------------------------------------------------------- #
# near-tie accounting used by the parity tests
# --------------------------------------------------------------------------- #
def top2_relative_gap(d: np.ndarray) -> np.ndarray:
    """Relative gap between the two smallest distances of each row (north_star near-tie rule)."""
    part = np.partition(d.astype(F64), 1, axis=-1)[:, :2]
    lo, hi = part[:, 0], part[:, 1]
    denom = np.maximum(np.abs(hi), np.finfo(F64).tiny)
    return (hi - lo) / denom


def classify_code_mismatches(z: np.ndarray, p: RqvaeParams, codes_ours: np.ndarray,
                             rel_tol: float = 1e-5) -> Tuple[int, int]:
    """Compare argmin codes with the oracle's on latents ``z``.

    Returns (near_tie_mismatches, hard_mismatches).  A differing row counts as a near-tie when,
    at the first level where the codes differ, the oracle's top-2 distance gap is below
    ``rel_tol`` relative (deeper levels of that row are then expected to differ as well).
    """
    resids, dists, codes_ref = rq_trace(z, p)
    near = hard = 0
    bad = np.nonzero((codes_ref != codes_ours).any(axis=1))[0]
    for i in bad:
        lvl = int(np.nonzero(codes_ref[i] != codes_ours[i])[0][0])
        gap = top2_relative_gap(dists[lvl][i:i + 1])[0]
        # also accept when our pick is within rel_tol of the oracle's minimum
        drow = dists[lvl][i].astype(F64)
        alt = abs(drow[codes_ours[i, lvl]] - drow.min()) / max(abs(drow[codes_ours[i, lvl]]), 1e-300)
        if gap < rel_tol or alt < rel_tol:
            near += 1
        else:
            hard += 1
    return near, hard
