"""Item-sharded index generation over the GPUs of one box (one process per GPU, NCCL over NVLink).

The reference's index/ directory is single-process; this is the north-star sharding of
``index/generate_indices.py:85-128``:

* PASS 0 shards naturally - contiguous item blocks per rank, encoder weights and codebooks
  replicated, no communication (the 16 KB/item embeddings never leave their rank);
* collision detection needs the global code table: after PASS 0 the packed codes (32 B/item) and
  the residuals entering the last level (128 B/item) are all-gathered ONCE;
* every round each rank derives the same CSR of collision groups from the same table
  (deterministic sort), resolves the groups ``g % world == rank`` with the per-group Sinkhorn
  kernel, and the last-level code deltas are summed with one all-reduce (8 B/item).  Groups are
  disjoint, so the result is identical to the single-GPU run bit for bit.

The arithmetic lives behind a small backend interface so that the host logic above can be
exercised with ``gloo`` on CPU in the tests (where the backend is the numpy oracle); the product
backend is ``CudaBackend`` (C-ABI kernels).  Nothing here falls back to the CPU by itself.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import torch
import torch.distributed as dist


@dataclass
class ShardPlan:
    """Contiguous block partition of ``n_total`` items over ``world`` ranks (first ranks get the remainder)."""
    n_total: int
    world: int

    def count(self, rank: int) -> int:
        base, rem = divmod(self.n_total, self.world)
        return base + (1 if rank < rem else 0)

    def start(self, rank: int) -> int:
        base, rem = divmod(self.n_total, self.world)
        return rank * base + min(rank, rem)

    def slice(self, rank: int) -> slice:
        return slice(self.start(rank), self.start(rank) + self.count(rank))

    @property
    def max_count(self) -> int:
        return self.count(0)


def all_gather_rows(local: torch.Tensor, plan: ShardPlan, rank: int, group=None) -> torch.Tensor:
    """Concatenate the ranks' row blocks (ragged blocks are padded to the largest for the collective)."""
    pad = plan.max_count
    buf = local
    if local.shape[0] != pad:
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    out = torch.empty((plan.world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if plan.n_total == plan.world * pad:
        return out
    return torch.cat([out[r * pad: r * pad + plan.count(r)] for r in range(plan.world)], dim=0)


class CudaBackend:
    """Kernels of liblcrec_b200.so on this rank's GPU."""

    def __init__(self, model, n_local_max: int, chunk_rows: int = 131072):
        from . import generate_indices as G
        from . import ops
        self.ops = ops
        self.model = model
        self.indexer = G.build_indexer(model, max(n_local_max, 1), min(chunk_rows, max(n_local_max, 1)))
        self.cbs = [vq.embedding.weight.detach() for vq in model.rq.vq_layers]
        self.n_codes = [int(c.shape[0]) for c in self.cbs]
        self.eps = float(model.rq.vq_layers[-1].sk_epsilon)
        self.iters = int(model.rq.vq_layers[-1].sk_iters)

    def pass0(self, x_local: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        n = x_local.shape[0]
        if n == 0:
            dev = self.cbs[0].device
            return (torch.zeros((0, len(self.cbs)), dtype=torch.int64, device=dev),
                    torch.zeros((0, self.cbs[0].shape[1]), dtype=torch.float32, device=dev))
        self.indexer.pass0(x_local, 0)
        return self.indexer.codes_view(n).clone(), self.indexer.resid_view(n).clone()

    def collisions(self, codes_all: torch.Tensor) -> dict:
        return self.ops.collisions(codes_all, self.n_codes)

    def resolve(self, resid_all: torch.Tensor, codes_all: torch.Tensor, info: dict, mod: int, rem: int) -> None:
        flags = self.ops.sinkhorn_groups(resid_all, self.cbs[-1], info["offsets"], info["members"],
                                         info["counts_dev"][1:2], info["n_groups"], info["n_rows"], self.eps,
                                         self.iters, codes_all, codes_all.shape[1] - 1, part_mod=mod, part_rem=rem)
        if flags & 6:
            raise RuntimeError(f"sinkhorn_groups failed with flags {flags}")


def generate_codes_sharded(backend, x_local: torch.Tensor, plan: ShardPlan, rank: int, max_rounds: int = 20,
                           group=None):
    """Sharded ``generate_indices``: returns (codes of this rank's items, stats).  Collective: every
    rank of ``group`` must call it."""
    codes_local, resid_local = backend.pass0(x_local)
    if plan.world > 1:
        codes_all = all_gather_rows(codes_local, plan, rank, group)
        resid_all = all_gather_rows(resid_local, plan, rank, group)
    else:
        codes_all, resid_all = codes_local, resid_local
    codes_all = codes_all.contiguous()
    n = plan.n_total
    rounds = 0
    first = None
    rows_total = 0
    while True:                                               # generate_indices.py:108-128
        info = backend.collisions(codes_all)
        if first is None:
            first = (info["n_groups"], info["n_rows"])
        if info["n_unique"] == n or rounds >= max_rounds:
            break
        last = codes_all.shape[1] - 1
        if plan.world > 1:
            old = codes_all[:, last].clone()
            backend.resolve(resid_all, codes_all, info, plan.world, rank)
            delta = codes_all[:, last] - old
            dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
            codes_all[:, last] = old + delta
        else:
            backend.resolve(resid_all, codes_all, info, 1, 0)
        rows_total += info["n_rows"]
        rounds += 1
    stats = {"rounds": rounds, "n_unique": info["n_unique"], "groups_round1": first[0], "rows_round1": first[1],
             "sinkhorn_rows": rows_total, "max_multiplicity": info["max_multiplicity"],
             "collision_rate": (n - info["n_unique"]) / max(n, 1)}
    return codes_all[plan.slice(rank)], stats
