#!/usr/bin/env python
"""Time the UNMODIFIED reference training loop ``Trainer._train_epoch`` (baseline/_ref/index/trainer.py:98-125) on CPU (or
``--device cuda:0``) for BASELINE configs[1]: run.sh architecture, batch 1024, AdamW lr 1e-3 wd 1e-4, linear warm-up, clip 1.0,
Sinkhorn on the last level.  Inputs come in an ``.npz`` (x, encoder / decoder Linear tensors, codebooks) written by bench.py;
nothing of lcrec_b200 is imported.  Prints one JSON object."""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_INDEX = os.path.join(HERE, "_ref", "index")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npz", required=True)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if not os.path.isfile(os.path.join(REF_INDEX, "trainer.py")):
        print(json.dumps({"unavailable": "baseline/_ref/index missing"}))
        return 0
    sys.path.insert(0, REF_INDEX)
    import numpy as np
    import torch
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    from models.rqvae import RQVAE
    from trainer import Trainer
    z = np.load(a.npz)
    x = torch.from_numpy(np.ascontiguousarray(z["x"], dtype=np.float32))
    n_lin = len([k for k in z.files if k.startswith("w") and not k.startswith("wd")])
    dims = [x.shape[1]] + [int(z[f"w{i}"].shape[0]) for i in range(n_lin)]
    cbs = [z[f"cb{l}"] for l in range(len([k for k in z.files if k.startswith("cb")]))]
    tmp = tempfile.mkdtemp(prefix="lcrec_reftrain_")
    args = argparse.Namespace(lr=1e-3, epochs=a.repeat, batch_size=a.batch, num_workers=0, eval_step=10 ** 9, learner="AdamW",
                              lr_scheduler_type="linear", warmup_epochs=1, data_path="", weight_decay=1e-4, dropout_prob=0.0,
                              bn=bool(a.bn), loss_type="mse", kmeans_init=False, kmeans_iters=100, sk_epsilons=[0.0, 0.0, 0.0, 0.003],
                              sk_iters=50, device=a.device, num_emb_list=[int(c.shape[0]) for c in cbs], e_dim=int(cbs[0].shape[1]),
                              quant_loss_weight=1.0, beta=0.25, layers=dims[1:-1], save_limit=5, ckpt_dir=tmp)
    model = RQVAE(in_dim=dims[0], num_emb_list=args.num_emb_list, e_dim=args.e_dim, layers=args.layers, bn=args.bn, kmeans_init=False,
                  sk_epsilons=args.sk_epsilons, sk_iters=args.sk_iters)
    sd = model.state_dict()
    stride = 4 if a.bn else 3
    for i in range(n_lin):
        sd[f"encoder.mlp_layers.{1 + stride * i}.weight"] = torch.from_numpy(z[f"w{i}"])
        sd[f"encoder.mlp_layers.{1 + stride * i}.bias"] = torch.from_numpy(z[f"b{i}"])
        sd[f"decoder.mlp_layers.{1 + stride * i}.weight"] = torch.from_numpy(z[f"wd{i}"])
        sd[f"decoder.mlp_layers.{1 + stride * i}.bias"] = torch.from_numpy(z[f"bd{i}"])
    for l, cb in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
    model.load_state_dict(sd)
    loader = torch.utils.data.DataLoader(x, batch_size=a.batch, shuffle=False)
    tr = Trainer(args, model, len(loader))
    secs, losses = [], []
    for ep in range(max(a.repeat, 1)):
        sink = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            losses.append(tr._train_epoch(loader, ep))
            if a.device.startswith("cuda"):
                torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
    res = {"seconds": secs, "items": int(x.shape[0]), "batches": len(loader), "losses": [[float(v) for v in l] for l in losses],
           "device": a.device, "torch_threads": torch.get_num_threads(), "cpu_count": os.cpu_count(), "torch": torch.__version__}
    s = json.dumps(res)
    if a.out:
        open(a.out, "w").write(s)
    print(s)
    return 0


if __name__ == "__main__":
    sys.exit(main())
