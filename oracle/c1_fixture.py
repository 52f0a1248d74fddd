"""Reader of tests/golden/c1_loop{,_bn}.npz (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

The fixture (oracle/make_golden_c1.py) holds what the UNMODIFIED reference produced for BASELINE configs[0] - 25 000
items, run.sh architecture: PASS-0 table, the table after every collision round (sparse row updates), final table, sha256
of the JSON it wrote, the k-means codebooks / BatchNorm tensors its own code produced.  Items and Linear weights are
regenerated here from their seeds (numpy PCG64 streams, ``lcrec_b200.synth``)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from lcrec_b200.synth import seeded_weights, synth_items

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


@dataclass
class C1Case:
    x: np.ndarray
    state_dict: Dict[str, np.ndarray]        # reference key layout (encoder Linear / BatchNorm, rq codebooks)
    tables: List[np.ndarray]                 # [PASS 0, after round 1, ..., after the last round], int64 (n, L)
    final: np.ndarray
    bn: bool
    dims: List[int]
    raw: Dict[str, np.ndarray]


def load_c1(name: str = "c1_loop", n_items: int | None = None) -> C1Case:
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    n, dims, bn = int(g["n"]), g["dims"].tolist(), bool(g["bn"])
    ws, bs, _ = seeded_weights(dims, [256] * 4, dims[-1], seed=int(g["seed_w"]))
    x = synth_items(n, dims[0], n_parents=n // 8, seed=int(g["seed_x"]))
    stride = 4 if bn else 3                  # Sequential: Dropout, Linear, [BatchNorm1d], ReLU (layers.py:22-30)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    for i, (w, b) in enumerate(zip(ws, bs)):
        sd[f"encoder.mlp_layers.{1 + stride * i}.weight"] = w
        sd[f"encoder.mlp_layers.{1 + stride * i}.bias"] = b
    tables = [g["codes_pass0"].astype(np.int64)]
    for r in range(1, int(g["n_rounds"]) + 1):
        t = tables[-1].copy()
        t[g[f"round{r}_items"]] = g[f"round{r}_codes"].astype(np.int64)
        tables.append(t)
    final = g["codes_final"].astype(np.int64)
    assert (tables[-1] == final).all()
    if n_items is not None:
        x = x[:n_items]
    return C1Case(x=x, state_dict=sd, tables=tables, final=final, bn=bn, dims=dims, raw=g)
