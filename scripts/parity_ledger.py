"""Collision-loop parity ledger on a B200 (diagnostic; the collected version is tests/test_gpu_loop_ledger.py).

Separates the LOOP from the ENCODER:
  (B) oracle loop started from the GPU's own PASS-0 table + last-level residuals  vs  the GPU loop
      -> must be identical row for row, every round, every kernel mode;
  (A) oracle loop from the oracle's own (numpy fp32) latents vs the GPU loop
      -> differences start from a few PRIMARY rows (same group, same incoming table, other pick) whose
         witness is the ulp-level difference of the two fp32 distance matrices; everything else is cascade.
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from lcrec_b200 import generate_indices as G
from lcrec_b200 import ops
from lcrec_b200.models import RQVAE
from lcrec_b200.synth import seeded_weights, synth_items
from oracle import lcrec_oracle as O

DEV = torch.device("cuda:0")
DIMS = [4096, 2048, 1024, 512, 256, 128, 64, 32]


def model_for(ws, bs, cbs):
    m = RQVAE(in_dim=4096, num_emb_list=[c.shape[0] for c in cbs], e_dim=cbs[0].shape[1], layers=DIMS[1:-1],
              sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
    sd = m.state_dict()
    lin = sorted([k for k in sd if k.startswith("encoder.mlp_layers.") and k.endswith(".weight")], key=lambda s: int(s.split(".")[2]))
    for k, w, b in zip(lin, ws, bs):
        sd[k] = torch.from_numpy(w); sd[k.replace(".weight", ".bias")] = torch.from_numpy(b)
    for l, cb in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
    m.load_state_dict(sd)
    return m.to(DEV).eval()


def case(n, n_parents, wseed, xseed, cb_scale):
    ws, bs, cbs = seeded_weights(DIMS, [256] * 4, 32, seed=wseed, cb_scale=cb_scale)
    x = synth_items(n, 4096, n_parents=n_parents, seed=xseed)
    m = model_for(ws, bs, cbs)
    xd = torch.from_numpy(x).to(DEV)
    p = O.RqvaeParams(encoder=O.MlpParams(ws, bs), codebooks=cbs, sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
    out = {"n": n}
    ix = G.build_indexer(m, n)
    ix.pass0(xd)
    c0 = ix.codes_view(n).cpu().numpy().copy()
    r3 = ix.resid_view(n).cpu().numpy().copy()
    t = time.time()
    fo, tro = O.resolve_collisions(c0, r3, cbs[3], 0.003, 50)
    out["oracle_loop_s"] = round(time.time() - t, 1)
    tabs_o = [c0] + tro.rounds
    # (B1) GPU free-running round by round (global sort + default Sinkhorn mode)
    for mode in (2, 0, 1):
        ops.sinkhorn_set_mode(mode)
        ix.codes_view(n).copy_(torch.from_numpy(c0).to(DEV))
        per = []
        for r in range(1, len(tabs_o)):
            ix.round(n)
            per.append(int((ix.codes_view(n).cpu().numpy() != tabs_o[r]).any(axis=1).sum()))
        out[f"B_rounds_mode{mode}"] = per
    ops.sinkhorn_set_mode(2)
    # (B2) the product entry points
    for seg in (1, 0):
        ops.indexer_set_segments(bool(seg))
        cd, st = ix.run_device(xd, 20)
        out[f"B_e2e_seg{seg}"] = int((cd.cpu().numpy() != fo).any(axis=1).sum())
        out[f"rounds_seg{seg}"] = st["rounds"]
    ops.indexer_set_segments(True)
    cd, st = ix.run_device(xd, 20)
    gpu_final = cd.cpu().numpy()
    # (A) oracle on its own latents
    z = O.mlp_forward(x, p.encoder)
    rs, _, cz = O.rq_trace(z, p)
    near, hard = O.classify_code_mismatches(z, p, c0)
    fz, trz = O.resolve_collisions(cz, rs[3], cbs[3], 0.003, 50)
    led = O.loop_ledger([cz] + trz.rounds, tabs_o)
    wit = []
    for r, i in led.primary:
        prev = ([cz] + trz.rounds)[r - 1]
        g = [g for g in O.collision_groups(prev) if i in g][0]
        u, rel = O.distance_ulp_witness(rs[3][g], r3[g], cbs[3])
        same_on_gpu_inputs = bool((O.vq_assign(r3[g], cbs[3], True, 0.003, 50) == tabs_o[r][g, 3]).all())
        wit.append({"round": r, "item": i, "group_size": len(g), "dist_ulp": u, "resid_rel": rel, "oracle_on_gpu_rows_agrees": same_on_gpu_inputs})
    out.update({"A_pass0_near": near, "A_pass0_hard": hard, "A_final_rows": led.rows_differing_final,
                "A_first_round": led.first_round, "A_primary": wit, "A_cascade": led.cascade, "A_per_round": led.per_round,
                "latent_rel_err_max": float(np.abs(rs[3] - r3).max() / np.abs(rs[3]).max()),
                "gpu_final_vs_oracle_own": int((gpu_final != fz).any(axis=1).sum())})
    return out


if __name__ == "__main__":
    res = [case(3000, 375, 3, 4, 0.02), case(20000, 2500, 3, 11, 0.02)]
    for r in res:
        print(json.dumps(r))
