"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/index).

Run in the build container only (``python oracle/make_golden.py``); /root/reference does not
exist on the GPU box, so tests read the committed fixtures, never the reference.

The reference has no tests / golden vectors; these fixtures are outputs of
its own code (torch CPU) on seeded inputs:

  sinkhorn_kat.npz    layers.sinkhorn_algorithm + VectorQuantizer.center_distance_for_constraint
  small_model.npz     RQVAE trained 3 epochs by the reference Trainer; encoder latents,
                      get_indices, forward(use_sk=True) values, the complete generate_indices.py
                      script result (exec'd with only its hard-coded path/device lines replaced)
                      and a per-round trace of its collision loop
  bn_model.npz        same with bn=True (what ``run.sh --bn False`` really runs), eval mode
  fullshape.npz       run.sh shape 4096->...->32, 4x256 codes: seeded weights, latents + codes
  trainer_steps.npz   loss trajectory of reference Trainer._train_epoch (no k-means)
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import random
import sys
import tempfile

import numpy as np

REF = "/root/reference/index"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

sys.path.insert(0, REF)          # bare module names: datasets, utils, models, trainer
import torch                      # noqa: E402
from models.rqvae import RQVAE    # noqa: E402
from models.layers import sinkhorn_algorithm   # noqa: E402
from models.vq import VectorQuantizer           # noqa: E402
from trainer import Trainer       # noqa: E402
from datasets import EmbDataset   # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from synth import seeded_weights, synth_items   # noqa: E402  (shared with the tests)


def seed_all(s=2024):
    random.seed(s); np.random.seed(s); torch.manual_seed(s)   # index/main.py:52-60


def sd_to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items() if v.ndim > 0 or "num_batches" not in k}


def make_args(**kw):
    base = dict(lr=1e-3, epochs=3, batch_size=256, num_workers=0, eval_step=50, learner="AdamW",
                lr_scheduler_type="linear", warmup_epochs=1, data_path="", weight_decay=1e-4,
                dropout_prob=0.0, bn=False, loss_type="mse", kmeans_init=True, kmeans_iters=10,
                sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50, device="cpu",
                num_emb_list=[32, 32, 32, 32], e_dim=16, quant_loss_weight=1.0, beta=0.25,
                layers=[64, 48], save_limit=5, ckpt_dir="")
    base.update(kw)
    return argparse.Namespace(**base)


def build_model(args, in_dim):
    return RQVAE(in_dim=in_dim, num_emb_list=args.num_emb_list, e_dim=args.e_dim, layers=args.layers,
                 dropout_prob=args.dropout_prob, bn=args.bn, loss_type=args.loss_type,
                 quant_loss_weight=args.quant_loss_weight, beta=args.beta, kmeans_init=args.kmeans_init,
                 kmeans_iters=args.kmeans_iters, sk_epsilons=args.sk_epsilons, sk_iters=args.sk_iters)


def run_generate_script(ckpt_path, out_dir):
    """exec index/generate_indices.py with ONLY lines 44-51 substituted (paths, device, and
    weights_only=False which torch>=2.6 needs to unpickle the argparse.Namespace)."""
    src = open(os.path.join(REF, "generate_indices.py")).read()
    subs = {
        'ckpt_path = "/zhengbowen/rqvae_ckpt/xxxx"': f'ckpt_path = {ckpt_path!r}',
        'output_dir = f"/zhengbowen/data/{dataset}/"': f'output_dir = {out_dir!r}',
        'device = torch.device("cuda:0")': 'device = torch.device("cpu")',
        "ckpt = torch.load(ckpt_path, map_location=torch.device('cpu'))":
            "ckpt = torch.load(ckpt_path, map_location=torch.device('cpu'), weights_only=False)",
    }
    for a, b in subs.items():
        assert a in src, a
        src = src.replace(a, b)
    g = {"__name__": "__ref_generate__"}
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        exec(compile(src, "generate_indices.py", "exec"), g)
    return g


def token_table_to_codes(all_indices):
    return np.array([[int(t[3:-1]) for t in row] for row in all_indices.tolist()], dtype=np.int64)


def traced_collision_loop(model, data_np, codes0, max_rounds=20):
    """The loop of generate_indices.py:107-128 driven step by step on the reference model so every
    round's table can be stored (teacher forcing)."""
    codes = codes0.copy()
    rounds, qargs = [], []
    tt = 0
    while tt < max_rounds:
        strs = ["-".join(map(str, r)) for r in codes.tolist()]
        if len(set(strs)) == len(strs):
            break
        table = {}
        for i, s in enumerate(strs):
            table.setdefault(s, []).append(i)
        for g in [v for v in table.values() if len(v) > 1]:
            d = torch.from_numpy(data_np[g])
            codes[g] = model.get_indices(d, use_sk=True).view(-1, codes.shape[1]).numpy()
        rounds.append(codes.copy())
        tt += 1
    return rounds


def gen_sinkhorn(path):
    rng = np.random.default_rng(11)
    out = {}
    cases = [(2, 256, 0.003), (5, 256, 0.003), (24, 256, 0.003), (64, 32, 0.01), (300, 256, 0.003),
             (1024, 256, 0.003), (7, 100, 0.05), (3, 256, 0.003), (16, 256, 0.003)]
    meta = []
    for ci, (n, k, eps) in enumerate(cases):
        d32 = rng.standard_normal((n, k)).astype(np.float32) ** 2 * 3.0
        if ci == 7:      # exact duplicate rows: the exact-tie regime
            d32[1] = d32[0]; d32[2] = d32[0]
        if ci == 8:      # clustered rows (near-duplicates), like a real collision group
            base = d32[0].copy()
            d32 = (base[None] + 1e-3 * rng.standard_normal((n, k))).astype(np.float32)
        dc = VectorQuantizer.center_distance_for_constraint(torch.from_numpy(d32))
        q = sinkhorn_algorithm(dc.double(), eps, 50)
        out[f"d_{ci}"] = d32
        out[f"dc_{ci}"] = dc.numpy()
        out[f"q_{ci}"] = q.numpy() if n <= 64 else q.numpy()[:16]   # big cases: head rows only
        out[f"arg_{ci}"] = torch.argmax(q, dim=-1).numpy()
        meta.append([n, k, eps, 50])
    out["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(path, **out)


def gen_model_case(path, bn, n_items=2000, in_dim=96):
    seed_all()
    x = synth_items(n_items, in_dim, n_parents=n_items // 8, seed=5)
    tmp = tempfile.mkdtemp(prefix="lcrec_golden_")
    data_path = os.path.join(tmp, "Toy.emb.npy")
    np.save(data_path, x)
    args = make_args(bn=bn, data_path=data_path, ckpt_dir=os.path.join(tmp, "ckpt"), eval_step=3)
    data = EmbDataset(data_path)
    model = build_model(args, data.dim)
    loader = torch.utils.data.DataLoader(data, num_workers=0, batch_size=args.batch_size, shuffle=True)
    trainer = Trainer(args, model, len(loader))
    with contextlib.redirect_stderr(io.StringIO()):
        best_loss, best_coll = trainer.fit(loader)
    ckpt = os.path.join(trainer.ckpt_dir, "best_collision_model.pth")
    model.eval()
    out = {"x": x, "best_loss": np.float64(best_loss), "best_collision_rate": np.float64(best_coll),
           "sk_epsilons": np.array(args.sk_epsilons), "sk_iters": np.int64(args.sk_iters),
           "num_emb_list": np.array(args.num_emb_list), "e_dim": np.int64(args.e_dim),
           "layers": np.array(args.layers), "bn": np.bool_(bn)}
    sd = torch.load(ckpt, map_location="cpu", weights_only=False)["state_dict"]
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    model.load_state_dict(sd)
    xt = torch.from_numpy(x)
    with torch.no_grad():
        z = model.encoder(xt)
        out["latents"] = z.numpy()
        out["codes_argmin_full"] = model.get_indices(xt, use_sk=False).numpy()
        xq, rq_loss, idx = model.rq(z, use_sk=False)
        out["rq_xq"] = xq.numpy(); out["rq_loss"] = rq_loss.numpy()
        b = xt[:512]
        o, l, i = model(b, use_sk=True)
        tot, rec = model.compute_loss(o, l, xs=b)
        out["fwd_out"] = o.numpy(); out["fwd_rq_loss"] = l.numpy(); out["fwd_idx"] = i.numpy()
        out["fwd_total"] = tot.numpy(); out["fwd_recon"] = rec.numpy()
    g = run_generate_script(ckpt, tmp + "/")
    codes_final = token_table_to_codes(g["all_indices"])
    out["script_codes_final"] = codes_final
    out["script_json"] = np.frombuffer(open(os.path.join(tmp, "Games.index.json"), "rb").read(), dtype=np.uint8)
    # per-round trace on a fresh model with the script's epsilon rule
    m2 = build_model(args, data.dim); m2.load_state_dict(sd); m2.eval()
    with torch.no_grad():
        codes0 = torch.cat([m2.get_indices(xt[s:s + 64], use_sk=False) for s in range(0, n_items, 64)]).numpy()
        for vq in m2.rq.vq_layers[:-1]:
            vq.sk_epsilon = 0.0
        if m2.rq.vq_layers[-1].sk_epsilon == 0.0:
            m2.rq.vq_layers[-1].sk_epsilon = 0.003
        rounds = traced_collision_loop(m2, x, codes0)
    out["codes_pass0"] = codes0
    out["rounds"] = np.stack(rounds) if rounds else np.zeros((0,) + codes0.shape, np.int64)
    assert (rounds[-1] == codes_final).all() if rounds else (codes0 == codes_final).all()
    np.savez_compressed(path, **out)
    print(os.path.basename(path), "collision rate pass0",
          1 - len({tuple(r) for r in codes0.tolist()}) / n_items, "rounds", len(rounds),
          "final", 1 - len({tuple(r) for r in codes_final.tolist()}) / n_items)


def gen_fullshape(path, n_items=512):
    dims = [4096, 2048, 1024, 512, 256, 128, 64, 32]
    ws, bs, cbs = seeded_weights(dims, [256] * 4, 32, seed=77)
    args = make_args(layers=dims[1:-1], e_dim=32, num_emb_list=[256] * 4, kmeans_init=False)
    model = build_model(args, 4096)
    sd = model.state_dict()
    lin = [k for k in sd if k.startswith("encoder.mlp_layers.") and k.endswith(".weight")]
    for k, w, b in zip(sorted(lin, key=lambda s: int(s.split(".")[2])), ws, bs):
        sd[k] = torch.from_numpy(w); sd[k.replace(".weight", ".bias")] = torch.from_numpy(b)
    for l, cb in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
    model.load_state_dict(sd); model.eval()
    x = synth_items(n_items, 4096, n_parents=n_items // 8, seed=9)
    with torch.no_grad():
        xt = torch.from_numpy(x)
        z = model.encoder(xt)
        codes = model.get_indices(xt, use_sk=False)
        h1 = torch.relu(torch.nn.functional.linear(xt, sd["encoder.mlp_layers.1.weight"], sd["encoder.mlp_layers.1.bias"]))
    np.savez_compressed(path, latents=z.numpy(), codes=codes.numpy(), h1_head=h1[:8, :64].numpy(),
                        dims=np.array(dims), seed_w=np.int64(77), seed_x=np.int64(9), n=np.int64(n_items))


def gen_trainer(path):
    seed_all()
    x = synth_items(1024, 96, n_parents=128, seed=6)
    tmp = tempfile.mkdtemp(prefix="lcrec_golden_")
    data_path = os.path.join(tmp, "Toy.emb.npy"); np.save(data_path, x)
    args = make_args(kmeans_init=False, data_path=data_path, ckpt_dir=os.path.join(tmp, "ckpt"), epochs=4)
    model = build_model(args, 96)
    init_sd = {k: v.clone() for k, v in model.state_dict().items()}
    loader = torch.utils.data.DataLoader(EmbDataset(data_path), batch_size=256, shuffle=False)
    tr = Trainer(args, model, len(loader))
    losses = []
    with contextlib.redirect_stderr(io.StringIO()):
        for ep in range(4):
            losses.append(tr._train_epoch(loader, ep))
        coll = tr._valid_epoch(loader)
    out = {"x": x, "losses": np.array(losses, dtype=np.float64), "collision_rate": np.float64(coll)}
    for k, v in init_sd.items():
        out["init/" + k] = v.numpy()
    for k, v in model.state_dict().items():
        out["final/" + k] = v.numpy()
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["sinkhorn", "small", "bn", "full", "trainer"]
    if "sinkhorn" in which: gen_sinkhorn(os.path.join(OUT, "sinkhorn_kat.npz"))
    if "small" in which: gen_model_case(os.path.join(OUT, "small_model.npz"), bn=False)
    if "bn" in which: gen_model_case(os.path.join(OUT, "bn_model.npz"), bn=True)
    if "full" in which: gen_fullshape(os.path.join(OUT, "fullshape.npz"))
    if "trainer" in which: gen_trainer(os.path.join(OUT, "trainer_steps.npz"))
    print("golden written to", os.path.abspath(OUT))
