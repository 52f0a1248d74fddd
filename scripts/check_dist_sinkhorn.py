"""torchrun --nproc-per-node N scripts/check_dist_sinkhorn.py : the row-sharded Sinkhorn (in-kernel peer-memory
all-reduce of the column marginals) against the single-GPU kernel on the concatenated batch.  Prints one JSON line."""
import json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops
from lcrec_b200.distributed import DistributedSinkhorn, ShardPlan

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
dist.barrier(); torch.cuda.synchronize()
os.dup2(saved, 1); os.close(saved)
out = {"world": world, "cases": []}
ok = True
for B, K, eps, iters in ((1024, 256, 0.003, 50), (1000, 256, 0.003, 50), (37, 64, 0.05, 20), (4096, 512, 0.01, 30)):
    g = torch.Generator(device=dev).manual_seed(1234 + B)
    raw = torch.rand(B, K, device=dev, generator=g) * 3.0           # identical on every rank
    sk = DistributedSinkhorn(K, dev)
    plan = ShardPlan(B, world)
    mine = raw[plan.slice(rank)]
    dc = sk.center(mine)
    ref_c = ops.center_distances(raw)
    cen_ok = bool(torch.equal(dc.double(), ref_c[plan.slice(rank)].double()))
    q, arg, flags = sk(dc, eps, iters, B)
    q_ref, arg_ref, _ = ops.sinkhorn_dense(ref_c, eps, iters, want_argmax=True)
    qr, ar = q_ref[plan.slice(rank)], arg_ref[plan.slice(rank)]
    rel = float(((q - qr).abs() / qr.abs().clamp_min(1e-300)).max()) if q.numel() else 0.0
    same = int((arg != ar).sum())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): sk(dc, eps, iters, B)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    res = torch.tensor([rel, float(same), float(int(flags.item())), 0.0 if cen_ok else 1.0, ms], dtype=torch.float64, device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    rel, same, fl, cen_bad, ms = res.tolist()
    out["cases"].append({"B": B, "K": K, "eps": eps, "iters": iters, "max_rel_err_Q": rel, "argmax_mismatches": int(same),
                         "flags": int(fl), "centring_exact": cen_bad == 0.0, "ms_per_call": round(ms, 3)})
    ok = ok and rel < 1e-9 and same == 0 and fl == 0 and cen_bad == 0.0
out["ok"] = ok
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
