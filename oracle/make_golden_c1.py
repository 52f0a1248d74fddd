"""tests/golden/c1_loop{,_bn}.npz: the collision loop of the UNMODIFIED reference at the BASELINE shape.

BASELINE.json configs[0] (C1): ``index/generate_indices.py`` on CPU, 25 000 synthetic 4096-d items, encoder
4096->2048->1024->512->256->128->64->32 (``index/run.sh``), 4 levels x 256 codes, e_dim 32, Sinkhorn (eps 0.003,
50 iterations) on the last level.  Run in the build container only (``python oracle/make_golden_c1.py``):

  * inputs: ``synth_items`` / ``seeded_weights`` (numpy PCG64 streams: regenerated from the seeds by the tests, not
    stored); BatchNorm parameters / running statistics (bn variant) and the codebooks come from the reference's OWN
    code - ten ``model.train()`` forward passes under ``no_grad`` (the first one runs ``VectorQuantizer.init_emb`` =
    scikit-learn k-means on each level's residuals, ``vq.py:40-49,67-68``; the others move the BN running statistics)
    - and ARE stored (128 KB + 64 KB);
  * the whole ``generate_indices.py`` script exec'd with only its hard-coded path / device lines substituted
    (``make_golden.run_generate_script``): final table + sha256 of the JSON bytes it wrote;
  * the same loop (``generate_indices.py:107-128``) driven step by step on the reference model so that the table after
    EVERY round is stored (first table + per-round sparse row updates), asserted equal to the script's final table;
  * noise floor: the numpy oracle (``generate_indices(reencode=True)``, OpenBLAS) against the reference (torch / MKL) -
    two IEEE fp32 CPU evaluations of the same algorithm - rows differing after every round.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG          # noqa: E402  (puts /root/reference/index on sys.path, imports the reference)
import torch                      # noqa: E402
import lcrec_oracle as O          # noqa: E402

DIMS = [4096, 2048, 1024, 512, 256, 128, 64, 32]
N_ITEMS, SEED_W, SEED_X = 25000, 21, 31


def build(bn: bool):
    from synth import seeded_weights, synth_items
    MG.seed_all()
    ws, bs, _ = seeded_weights(DIMS, [256] * 4, 32, seed=SEED_W)
    x = synth_items(N_ITEMS, 4096, n_parents=N_ITEMS // 8, seed=SEED_X)
    args = MG.make_args(layers=DIMS[1:-1], e_dim=32, num_emb_list=[256] * 4, kmeans_init=True, kmeans_iters=100, bn=bn,
                        batch_size=1024)
    model = MG.build_model(args, 4096)
    sd = model.state_dict()
    lin = sorted([k for k in sd if k.startswith("encoder.mlp_layers.") and k.endswith(".weight") and sd[k].ndim == 2],
                 key=lambda s: int(s.split(".")[2]))
    assert len(lin) == len(ws)
    for k, w, b in zip(lin, ws, bs):
        sd[k] = torch.from_numpy(w); sd[k.replace(".weight", ".bias")] = torch.from_numpy(b)
    model.load_state_dict(sd)
    model.train()
    with torch.no_grad():
        for s in range(10):
            model(torch.from_numpy(x[s * 1024:(s + 1) * 1024]))       # call 0: init_emb (k-means) on every level
    model.eval()
    return model, args, x


def gen(path: str, bn: bool):
    t0 = time.time()
    model, args, x = build(bn)
    tmp = tempfile.mkdtemp(prefix="lcrec_c1_")
    data_path = os.path.join(tmp, "C1.emb.npy")
    np.save(data_path, x)
    args.data_path = data_path
    ckpt = os.path.join(tmp, "c1.pth")
    torch.save({"args": args, "epoch": 0, "best_loss": 0.0, "best_collision_rate": 0.0, "state_dict": model.state_dict(),
                "optimizer": {}}, ckpt, pickle_protocol=4)
    g = MG.run_generate_script(ckpt, tmp + "/")
    final = MG.token_table_to_codes(g["all_indices"])
    js = open(os.path.join(tmp, "Games.index.json"), "rb").read()
    print(f"[{time.time() - t0:.0f}s] script done: collision rate {1 - O.n_unique_codes(final) / N_ITEMS:.5f}")
    # per-round trace on a fresh model with the script's epsilon rule (generate_indices.py:101-105)
    sd = model.state_dict()
    m2 = MG.build_model(args, 4096); m2.load_state_dict(sd); m2.eval()
    xt = torch.from_numpy(x)
    with torch.no_grad():
        codes0 = torch.cat([m2.get_indices(xt[s:s + 64], use_sk=False) for s in range(0, N_ITEMS, 64)]).numpy()
        for vq in m2.rq.vq_layers[:-1]:
            vq.sk_epsilon = 0.0
        if m2.rq.vq_layers[-1].sk_epsilon == 0.0:
            m2.rq.vq_layers[-1].sk_epsilon = 0.003
        rounds = MG.traced_collision_loop(m2, x, codes0)
    assert (rounds[-1] == final).all()
    print(f"[{time.time() - t0:.0f}s] traced loop: {len(rounds)} rounds")
    out = {"n": np.int64(N_ITEMS), "seed_w": np.int64(SEED_W), "seed_x": np.int64(SEED_X), "dims": np.array(DIMS),
           "bn": np.bool_(bn), "codes_pass0": codes0.astype(np.int16), "codes_final": final.astype(np.int16),
           "n_rounds": np.int64(len(rounds)), "json_sha256": np.frombuffer(hashlib.sha256(js).digest(), dtype=np.uint8),
           "json_len": np.int64(len(js))}
    prev = codes0
    for r, tab in enumerate(rounds):
        ch = np.nonzero((tab != prev).any(axis=1))[0]
        out[f"round{r + 1}_items"] = ch.astype(np.int32)
        out[f"round{r + 1}_codes"] = tab[ch].astype(np.int16)
        prev = tab
    for k, v in sd.items():
        if k.startswith("rq.") or (bn and v.ndim == 1 and int(k.split(".")[2]) % 4 == 2):
            out["sd/" + k] = v.numpy()
    # first latents / a few group latents: lets a test place the GPU encoder against the reference's numbers
    with torch.no_grad():
        out["latents_head"] = m2.encoder(xt[:512]).numpy()
    # noise floor between two IEEE fp32 CPU evaluations of the same script (numpy/OpenBLAS oracle vs torch/MKL reference)
    p = O.params_from_state_dict({k: v.numpy() for k, v in sd.items()}, [0.0, 0.0, 0.0, 0.003], 50)
    _, tr = O.generate_indices(x, p, batch_size=64, reencode=True)
    led = O.loop_ledger([codes0] + rounds, [tr.codes_pass0] + tr.rounds)
    out["oracle_vs_reference_rows_per_round"] = np.array(led.per_round, dtype=np.int64)
    out["oracle_vs_reference_primary"] = np.array(led.primary, dtype=np.int64).reshape(-1, 2)
    out["oracle_vs_reference_cascade"] = np.int64(led.cascade)
    print(f"[{time.time() - t0:.0f}s] oracle vs reference rows differing per round: {led.per_round}; "
          f"primary {led.primary}; cascade {led.cascade}")
    np.savez_compressed(path, **out)
    groups = [len(O.collision_groups(t)) for t in [codes0] + rounds[:-1]]
    print(os.path.basename(path), f"{os.path.getsize(path) / 1e6:.2f} MB; groups per round {groups}; "
          f"pass-0 collision rate {1 - O.n_unique_codes(codes0) / N_ITEMS:.4f}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["plain", "bn"]
    if "plain" in which: gen(os.path.join(MG.OUT, "c1_loop.npz"), bn=False)
    if "bn" in which: gen(os.path.join(MG.OUT, "c1_loop_bn.npz"), bn=True)
