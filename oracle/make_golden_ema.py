"""Generate tests/golden/ema_kat.npz by running the UNMODIFIED reference EMA quantiser
(/root/reference/index_improve/models/vq.py, torch CPU) on seeded inputs.  Build container only.

Recorded: a few training-mode forward calls of index_improve VectorQuantizer (EMA codebook update, vq.py:146-193) with
and without Sinkhorn: per step the latent batch, the indices it chose, the codebook / _ema_cluster_size / _ema_w after the
step, the loss and get_codebook_usage().  Cases 0/1 cross no reset step.  Case 2 has more codes than rows (codes whose smoothed
count stays below epsilon keep their vector) and reset_interval 3: `_reset_unused_codes` is wrapped (not changed) to record the
buffers it sees on entry; what it writes comes from torch's CPU RNG and is recorded only as "which rows were reset"."""
from __future__ import annotations

import os
import sys

import numpy as np

REF = "/root/reference/index_improve"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)
import torch                                   # noqa: E402
from models.vq import VectorQuantizer          # noqa: E402


def main():
    torch.manual_seed(2024)
    np.random.seed(2024)
    out = {}
    for case, (n_e, e_dim, use_sk, sk_eps, batch) in enumerate([(64, 16, False, 0.0, 300), (32, 8, True, 0.003, 96)]):  # noqa: E501
        vq = VectorQuantizer(n_e, e_dim, beta=0.25, kmeans_init=False, sk_epsilon=sk_eps, sk_iters=50, ema_decay=0.99,
                             epsilon=1e-5, reset_threshold=1e-5, reset_interval=1000)
        vq.train()
        out[f"c{case}_cfg"] = np.array([n_e, e_dim, int(use_sk), batch], dtype=np.int64)
        out[f"c{case}_sk_eps"] = np.array(sk_eps)
        out[f"c{case}_codebook0"] = vq.embedding.weight.detach().numpy().copy()
        steps = 6
        for s in range(steps):
            x = (torch.randn(batch, e_dim) * 0.02).float()
            x_q, loss, idx = vq(x, use_sk=use_sk, use_ema=True)
            out[f"c{case}_s{s}_x"] = x.numpy().copy()
            out[f"c{case}_s{s}_idx"] = idx.numpy().copy()
            out[f"c{case}_s{s}_loss"] = np.array(loss.item(), dtype=np.float32)
            out[f"c{case}_s{s}_codebook"] = vq.embedding.weight.detach().numpy().copy()
            out[f"c{case}_s{s}_cs"] = vq._ema_cluster_size.numpy().copy()
            out[f"c{case}_s{s}_w"] = vq._ema_w.numpy().copy()
            u = vq.get_codebook_usage()
            out[f"c{case}_s{s}_usage"] = np.array([u["utilization"], u["used_codes"], u["total_codes"]], dtype=np.float64)
        # use_ema=False leaves the codebook alone
        before = vq.embedding.weight.detach().numpy().copy()
        vq(x, use_sk=use_sk, use_ema=False)
        assert np.array_equal(before, vq.embedding.weight.detach().numpy())
        out[f"c{case}_steps"] = np.array(steps)
    # case 2: dead codes + reset
    case, n_e, e_dim, batch = 2, 128, 8, 40
    vq = VectorQuantizer(n_e, e_dim, beta=0.25, kmeans_init=False, sk_epsilon=0.0, sk_iters=50, ema_decay=0.99,
                         epsilon=1e-5, reset_threshold=1e-5, reset_interval=3)
    vq.train()
    seen = {}
    inner = vq._reset_unused_codes

    def recording_reset(latent):
        seen["cs"] = vq._ema_cluster_size.numpy().copy()
        seen["codebook"] = vq.embedding.weight.detach().numpy().copy()
        inner(latent)
    vq._reset_unused_codes = recording_reset
    out[f"c{case}_cfg"] = np.array([n_e, e_dim, 0, batch], dtype=np.int64)
    out[f"c{case}_sk_eps"] = np.array(0.0)
    out[f"c{case}_codebook0"] = vq.embedding.weight.detach().numpy().copy()
    steps = 4
    for s in range(steps):
        x = (torch.randn(batch, e_dim) * 0.02).float()
        if s < 2:
            x[:, 0] += 0.5                     # a corner of the codebook: most codes never win
        x_q, loss, idx = vq(x, use_sk=False, use_ema=True)
        out[f"c{case}_s{s}_x"] = x.numpy().copy()
        out[f"c{case}_s{s}_idx"] = idx.numpy().copy()
        out[f"c{case}_s{s}_loss"] = np.array(loss.item(), dtype=np.float32)
        out[f"c{case}_s{s}_codebook"] = vq.embedding.weight.detach().numpy().copy()
        out[f"c{case}_s{s}_cs"] = vq._ema_cluster_size.numpy().copy()
        out[f"c{case}_s{s}_w"] = vq._ema_w.numpy().copy()
        u = vq.get_codebook_usage()
        out[f"c{case}_s{s}_usage"] = np.array([u["utilization"], u["used_codes"], u["total_codes"]], dtype=np.float64)
        if "cs" in seen:
            out[f"c{case}_s{s}_pre_reset_cs"] = seen.pop("cs")
            out[f"c{case}_s{s}_pre_reset_codebook"] = seen.pop("codebook")
    out[f"c{case}_steps"] = np.array(steps)
    np.savez_compressed(os.path.join(OUT, "ema_kat.npz"), **out)
    print("wrote ema_kat.npz", {k: v.shape for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
