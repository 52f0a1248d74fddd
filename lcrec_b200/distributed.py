"""Item-sharded index generation over the GPUs of one box (one process per GPU, NCCL over NVLink).

The reference's index/ directory is single-process; this is the north-star sharding of
``index/generate_indices.py:85-128``:

* PASS 0 shards naturally - contiguous item blocks per rank, encoder weights and codebooks
  replicated, no communication (the 16 KB/item embeddings never leave their rank);
* collision detection needs a global view, but two items can only ever collide if they share the
  first L-1 codes (the rounds rewrite the LAST level only, generate_indices.py:101-105), so the
  global problem splits into independent PREFIX BUCKETS.  After PASS 0 the codes (32 B/item) and
  the residuals entering the last level (128 B/item) are all-gathered once; rank r then owns the
  buckets with ``prefix % world == r``, runs the whole <=20-round loop on them locally (1/world of
  the sort and Sinkhorn work, no per-round communication), and one all-reduce of the last-level
  codes (8 B/item) returns the results.  Groups, member order and arithmetic are exactly those of
  the single-GPU run, so the result is identical bit for bit.

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

    def collisions(self, codes: torch.Tensor) -> dict:
        return self.ops.collisions(codes, self.n_codes)

    def resolve_all(self, resid: torch.Tensor, codes: torch.Tensor, max_rounds: int):
        """All rounds inside the library (lcrec_indexer_resolve); None when the table exceeds the indexer's capacity."""
        if codes.shape[0] > self.indexer.max_items:
            return None
        return self.indexer.resolve_device(codes, resid.contiguous(), max_rounds)

    def resolve(self, resid: torch.Tensor, codes: torch.Tensor, info: dict) -> None:
        flags = self.ops.sinkhorn_groups(resid, self.cbs[-1], info["offsets"], info["members"],
                                         info["counts_dev"][1:2], info["n_groups"], info["n_rows"], self.eps,
                                         self.iters, codes, codes.shape[1] - 1)
        if flags & 6:
            raise RuntimeError(f"sinkhorn_groups failed with flags {flags}")


def bucket_owner(codes: torch.Tensor, n_codes, world: int) -> torch.Tensor:
    """Owner rank of every item's prefix bucket: the mixed-radix value of the first L-1 codes, mixed with a
    multiplicative hash (so that skewed code usage still spreads evenly), modulo world."""
    prefix = torch.zeros(codes.shape[0], dtype=torch.int64, device=codes.device)
    for l in range(codes.shape[1] - 1):
        prefix = prefix * int(n_codes[l]) + codes[:, l]
    mixed = (prefix * 0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF) >> 24      # wraps mod 2^64, then the high bits
    return mixed % world


def resolve_rounds(backend, codes: torch.Tensor, resid: torch.Tensor, max_rounds: int):
    """generate_indices.py:108-128 on one self-contained set of items (codes is updated in place)."""
    n = codes.shape[0]
    if n > 0 and hasattr(backend, "resolve_all"):
        st = backend.resolve_all(resid, codes, max_rounds)
        if st is not None:
            return {k: st[k] for k in ("rounds", "n_unique", "groups_round1", "rows_round1", "sinkhorn_rows", "max_multiplicity")}
    rounds = rows_total = 0
    first = (0, 0)
    info = {"n_unique": n, "max_multiplicity": 1 if n else 0}
    while n > 0:
        info = backend.collisions(codes)
        if rounds == 0:
            first = (info["n_groups"], info["n_rows"])
        if info["n_unique"] == n or rounds >= max_rounds:
            break
        backend.resolve(resid, codes, info)
        rows_total += info["n_rows"]
        rounds += 1
    return {"rounds": rounds, "n_unique": info["n_unique"], "groups_round1": first[0], "rows_round1": first[1],
            "sinkhorn_rows": rows_total, "max_multiplicity": info["max_multiplicity"]}


def generate_codes_sharded(backend, x_local: torch.Tensor, plan: ShardPlan, rank: int, max_rounds: int = 20,
                           group=None, n_codes=None):
    """Sharded ``generate_indices``: returns (codes of this rank's items, stats).  Collective: every
    rank of ``group`` must call it."""
    codes_local, resid_local = backend.pass0(x_local)
    n_codes = n_codes or getattr(backend, "n_codes")
    if plan.world == 1:
        codes = codes_local.contiguous()
        stats = resolve_rounds(backend, codes, resid_local, max_rounds)
        stats["collision_rate"] = (plan.n_total - stats["n_unique"]) / max(plan.n_total, 1)
        return codes, stats
    codes_all = all_gather_rows(codes_local, plan, rank, group)
    resid_all = all_gather_rows(resid_local, plan, rank, group)
    mine = torch.nonzero(bucket_owner(codes_all, n_codes, plan.world) == rank).squeeze(1)   # ascending item ids
    sub_codes = codes_all.index_select(0, mine).contiguous()
    sub_resid = resid_all.index_select(0, mine).contiguous()
    st = resolve_rounds(backend, sub_codes, sub_resid, max_rounds)
    last = codes_all.shape[1] - 1
    final_last = torch.zeros(plan.n_total, dtype=torch.int64, device=codes_all.device)
    final_last[mine] = sub_codes[:, last]
    agg = torch.tensor([st["n_unique"], st["groups_round1"], st["rows_round1"], st["sinkhorn_rows"]], dtype=torch.int64,
                       device=codes_all.device)
    mx = torch.tensor([st["rounds"], st["max_multiplicity"]], dtype=torch.int64, device=codes_all.device)
    dist.all_reduce(final_last, op=dist.ReduceOp.SUM, group=group)       # buckets are disjoint: a sum is a scatter
    dist.all_reduce(agg, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out = codes_local.clone()
    out[:, last] = final_last[plan.slice(rank)]
    n_unique, g1, r1, rows = [int(v) for v in agg.tolist()]
    rounds, max_mult = [int(v) for v in mx.tolist()]
    stats = {"rounds": rounds, "n_unique": n_unique, "groups_round1": g1, "rows_round1": r1, "sinkhorn_rows": rows,
             "max_multiplicity": max_mult, "collision_rate": (plan.n_total - n_unique) / max(plan.n_total, 1),
             "bucket_items_this_rank": int(mine.numel())}
    return out, stats
