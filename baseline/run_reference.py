#!/usr/bin/env python
"""Run the UNMODIFIED reference script ``index/generate_indices.py`` from ``baseline/_ref/index`` and time it.

``baseline/_ref/index`` is a verbatim, git-ignored copy of ``/root/reference/index`` staged by ``__graft_entry__.build()``
in the build container (the reference has no setup.py / pyproject.toml, so ``pip install --target baseline/_ref`` is not
possible: "Directory is not installable").  This runner is the only code of this repo on that path: it writes the inputs
the script reads (an ``.npy`` of embeddings and a ``.pth`` checkpoint in the reference's own format), substitutes the
script's hard-coded path / device lines (``generate_indices.py:44-51``, plus ``weights_only=False`` which torch >= 2.6 needs
to unpickle the ``argparse.Namespace``) and ``exec``s it.  Nothing of lcrec_b200 / liblcrec_b200.so is imported.

    python baseline/run_reference.py --npz inputs.npz --device cpu --threads 32 --out result.json

``inputs.npz``: ``x`` (n, in_dim) fp32, ``w{i}`` / ``b{i}`` encoder Linear tensors, ``cb{l}`` codebooks, ``eps`` (L,),
``sk_iters``.  Timed region = the exec of the whole script (checkpoint load, DataLoader PASS 0 at batch 64, <= 20
collision rounds with per-group re-encoding, JSON dump).  Prints one JSON object; ``--codes-out`` stores the final table.
"""
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
    ap.add_argument("--out", default=None)
    ap.add_argument("--codes-out", dest="codes_out", default=None)
    a = ap.parse_args()
    if not os.path.isfile(os.path.join(REF_INDEX, "generate_indices.py")):
        print(json.dumps({"unavailable": "baseline/_ref/index missing (staged by __graft_entry__.build() where /root/reference exists)"}))
        return 0
    sys.path.insert(0, REF_INDEX)            # the script imports bare module names: datasets, models (generate_indices.py:13-14)
    import numpy as np
    import torch
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    z = np.load(a.npz)
    x = np.ascontiguousarray(z["x"], dtype=np.float32)
    n_lin = len([k for k in z.files if k.startswith("w")])
    dims = [x.shape[1]] + [int(z[f"w{i}"].shape[0]) for i in range(n_lin)]
    cbs = [z[f"cb{l}"] for l in range(len([k for k in z.files if k.startswith("cb")]))]
    args = argparse.Namespace(data_path="", num_emb_list=[int(c.shape[0]) for c in cbs], e_dim=int(cbs[0].shape[1]), layers=dims[1:-1],
                              dropout_prob=0.0, bn=False, loss_type="mse", quant_loss_weight=1.0, kmeans_init=False, kmeans_iters=100,
                              sk_epsilons=[float(e) for e in z["eps"]], sk_iters=int(z["sk_iters"]), num_workers=0, beta=0.25)
    tmp = tempfile.mkdtemp(prefix="lcrec_ref_")
    args.data_path = os.path.join(tmp, "Bench.emb.npy")
    np.save(args.data_path, x)
    from models.rqvae import RQVAE
    model = RQVAE(in_dim=dims[0], num_emb_list=args.num_emb_list, e_dim=args.e_dim, layers=args.layers, sk_epsilons=args.sk_epsilons,
                  sk_iters=args.sk_iters)
    sd = model.state_dict()
    for i in range(n_lin):
        sd[f"encoder.mlp_layers.{1 + 3 * i}.weight"] = torch.from_numpy(z[f"w{i}"])
        sd[f"encoder.mlp_layers.{1 + 3 * i}.bias"] = torch.from_numpy(z[f"b{i}"])
    for l, cb in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
    model.load_state_dict(sd)
    ckpt = os.path.join(tmp, "bench.pth")
    torch.save({"args": args, "epoch": 0, "best_loss": 0.0, "best_collision_rate": 0.0, "state_dict": model.state_dict()}, ckpt,
               pickle_protocol=4)
    del model
    src = open(os.path.join(REF_INDEX, "generate_indices.py")).read()
    subs = {'ckpt_path = "/zhengbowen/rqvae_ckpt/xxxx"': f"ckpt_path = {ckpt!r}",
            'output_dir = f"/zhengbowen/data/{dataset}/"': f"output_dir = {tmp + '/'!r}",
            'device = torch.device("cuda:0")': f"device = torch.device({a.device!r})",
            "ckpt = torch.load(ckpt_path, map_location=torch.device('cpu'))":
                "ckpt = torch.load(ckpt_path, map_location=torch.device('cpu'), weights_only=False)"}
    for old, new in subs.items():
        assert old in src, old
        src = src.replace(old, new)
    code = compile(src, "generate_indices.py", "exec")
    secs, g = [], None
    for _ in range(max(a.repeat, 1)):
        g = {"__name__": "__ref_generate__"}
        sink = io.StringIO()                          # the script prints every collision group every round (:114)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            exec(code, g)
            if a.device.startswith("cuda"):
                torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
    toks = g["all_indices"]
    codes = np.array([[int(t[3:-1]) for t in row] for row in toks.tolist()], dtype=np.int64)
    n = codes.shape[0]
    uniq = len({tuple(r) for r in codes.tolist()})
    res = {"seconds": secs, "items": n, "items_per_s": n / (sum(secs) / len(secs)), "device": a.device, "rounds": int(g["tt"]),
           "collision_rate_final": (n - uniq) / n, "torch_threads": torch.get_num_threads(), "cpu_count": os.cpu_count(),
           "torch": torch.__version__}
    if a.codes_out:
        np.save(a.codes_out, codes)
    s = json.dumps(res)
    if a.out:
        open(a.out, "w").write(s)
    print(s)
    return 0


if __name__ == "__main__":
    sys.exit(main())
