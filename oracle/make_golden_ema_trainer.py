"""Generate tests/golden/ema_trainer_steps.npz: loss trajectory, utilisation and final buffers of the UNMODIFIED reference
``index_improve`` Trainer._train_epoch (torch CPU, 4 epochs x 4 batches of 256, AdamW + linear warm-up + clip 1.0, EMA codebook
step in every batch, Sinkhorn on the last level, no k-means, no reset step crossed).  Build container only.

Shim (not a behaviour change): ``index_improve/trainer.py:131`` asks ``torch.npu.is_available()`` (an Ascend build of
torch); stock torch has no ``torch.npu``, so a stub namespace answering False is attached before the run."""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import random
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/index_improve"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, REF)
import torch                                   # noqa: E402
if not hasattr(torch, "npu"):
    torch.npu = types.SimpleNamespace(is_available=lambda: False, empty_cache=lambda: None, manual_seed_all=lambda s: None)
from models.rqvae import RQVAE                 # noqa: E402
from trainer import Trainer                    # noqa: E402
from datasets import EmbDataset                # noqa: E402
sys.path.insert(0, HERE)
from synth import synth_items                  # noqa: E402


def main():
    random.seed(2024); np.random.seed(2024); torch.manual_seed(2024)
    torch.set_num_threads(8)
    x = synth_items(1024, 96, n_parents=128, seed=6)
    tmp = tempfile.mkdtemp(prefix="lcrec_golden_ema_")
    data_path = os.path.join(tmp, "Toy.emb.npy"); np.save(data_path, x)
    args = argparse.Namespace(lr=1e-3, epochs=4, batch_size=256, num_workers=0, eval_step=50, learner="AdamW",
                              lr_scheduler_type="linear", warmup_epochs=1, data_path=data_path, weight_decay=1e-4,
                              dropout_prob=0.0, bn=False, loss_type="mse", kmeans_init=False, kmeans_iters=10,
                              sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50, device="cpu", num_emb_list=[32] * 4, e_dim=16,
                              quant_loss_weight=1.0, beta=0.25, layers=[64, 48], save_limit=5,
                              ckpt_dir=os.path.join(tmp, "ckpt"), ema_decay=0.99, epsilon=1e-5, reset_threshold=1e-5,
                              reset_interval=1000)
    model = RQVAE(in_dim=96, num_emb_list=args.num_emb_list, e_dim=args.e_dim, layers=args.layers, dropout_prob=0.0, bn=False,
                  loss_type="mse", quant_loss_weight=1.0, beta=0.25, kmeans_init=False, kmeans_iters=10,
                  sk_epsilons=args.sk_epsilons, sk_iters=50, ema_decay=0.99, epsilon=1e-5, reset_threshold=1e-5,
                  reset_interval=1000)
    init_sd = {k: v.clone() for k, v in model.state_dict().items()}
    loader = torch.utils.data.DataLoader(EmbDataset(data_path), batch_size=256, shuffle=False)
    tr = Trainer(args, model, len(loader))
    losses = []
    with contextlib.redirect_stderr(io.StringIO()):
        for ep in range(4):
            losses.append(tr._train_epoch(loader, ep))
        coll = tr._valid_epoch(loader)
    avg_util, stats = tr._get_codebook_utilization()
    out = {"x": x, "losses": np.array(losses, dtype=np.float64), "collision_rate": np.float64(coll),
           "avg_utilization": np.float64(avg_util), "used_codes": np.array([s["used_codes"] for s in stats], dtype=np.int64)}
    for k, v in init_sd.items():
        out["init/" + k] = v.numpy()
    for k, v in model.state_dict().items():
        out["final/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "ema_trainer_steps.npz"), **out)
    print("losses", out["losses"], "collision", coll, "util", avg_util, out["used_codes"])


if __name__ == "__main__":
    main()
