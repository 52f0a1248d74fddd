"""torchrun --nproc-per-node N scripts/check_dp_trainer.py : data-parallel training (lcrec_b200.dp_trainer) against the loss
trajectory of the UNMODIFIED single-device reference Trainer (tests/golden/trainer_steps.npz: 4 epochs x 4 batches of 256,
AdamW + linear warm-up + clip 1.0, Sinkhorn on the last level) from the same initial state, and the step time of
BASELINE configs[1] (global batch 1024) on N GPUs.  Prints one JSON line."""
import argparse, json, os, sys, tempfile, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200.dp_trainer import DataParallelTrainer
from lcrec_b200.models import RQVAE

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = dict(np.load(os.path.join(root, "tests", "golden", "trainer_steps.npz")))


def make_args(**kw):
    base = dict(lr=1e-3, epochs=4, batch_size=256, num_workers=0, eval_step=50, learner="AdamW", lr_scheduler_type="linear",
                warmup_epochs=1, data_path="", weight_decay=1e-4, dropout_prob=0.0, bn=False, loss_type="mse", kmeans_init=False,
                kmeans_iters=10, sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50, device=str(dev), num_emb_list=[32] * 4, e_dim=16,
                quant_loss_weight=1.0, beta=0.25, layers=[64, 48], save_limit=5, ckpt_dir=tempfile.mkdtemp(prefix="dp_ckpt_"))
    base.update(kw)
    return argparse.Namespace(**base)


# ---- parity: 2e-3 on the reference's loss trajectory (the bar of the single-GPU trainer test)
args = make_args()
m = RQVAE(in_dim=96, num_emb_list=args.num_emb_list, e_dim=16, layers=args.layers, kmeans_init=False,
          sk_epsilons=args.sk_epsilons, sk_iters=50)
m.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init/")})
loader = torch.utils.data.DataLoader(torch.from_numpy(g["x"]), batch_size=256, shuffle=False)
tr = DataParallelTrainer(args, m, len(loader))
losses = np.array([tr._train_epoch(loader, ep) for ep in range(4)])
rel = float(np.abs(losses / g["losses"] - 1).max())
coll = tr._valid_epoch(loader)
sync = torch.cat([p.detach().reshape(-1)[:64].double() for p in m.parameters()])
lo, hi = sync.clone(), sync.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
replicas_equal = bool(torch.equal(lo, hi))

# ---- the run.sh shape against the UNMODIFIED reference Trainer's per-step losses (tests/golden/trainer_c2_steps.npz), bn False and
# True: under DP the BatchNorm kernels all-reduce their per-channel sums (synchronised BN = the single-device global batch)
from lcrec_b200.synth import seeded_weights, synth_items
g2 = dict(np.load(os.path.join(root, "tests", "golden", "trainer_c2_steps.npz")))
D = [4096, 2048, 1024, 512, 256, 128, 64, 32]
c2 = {}
for bn in (False, True):
    ws, bs, cbs = seeded_weights(D, [256] * 4, 32, seed=int(g2["seed_w"]), cb_scale=float(g2["cb_scale"]))
    wd, bd, _ = seeded_weights(D[::-1], [256] * 4, 32, seed=int(g2["seed_wd"]))
    mm = RQVAE(in_dim=4096, num_emb_list=[256] * 4, e_dim=32, layers=D[1:-1], bn=bn, kmeans_init=False, sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
    sd = mm.state_dict()
    stride = 4 if bn else 3
    for i, (w, b) in enumerate(zip(ws, bs)):
        sd[f"encoder.mlp_layers.{1 + stride * i}.weight"] = torch.from_numpy(w); sd[f"encoder.mlp_layers.{1 + stride * i}.bias"] = torch.from_numpy(b)
    for i, (w, b) in enumerate(zip(wd, bd)):
        sd[f"decoder.mlp_layers.{1 + stride * i}.weight"] = torch.from_numpy(w); sd[f"decoder.mlp_layers.{1 + stride * i}.bias"] = torch.from_numpy(b)
    for l, cb in enumerate(cbs):
        sd[f"rq.vq_layers.{l}.embedding.weight"] = torch.from_numpy(cb)
    mm.load_state_dict(sd)
    steps, batch = int(g2["steps"]), int(g2["batch"])
    xx = synth_items(steps * batch, 4096, n_parents=steps * batch // 8, seed=int(g2["seed_x"]))
    trc = DataParallelTrainer(make_args(layers=D[1:-1], num_emb_list=[256] * 4, e_dim=32, epochs=steps, warmup_epochs=2, bn=bn, batch_size=batch), mm, 1)
    got = np.array([trc._train_epoch([torch.from_numpy(xx[s * batch:(s + 1) * batch])], s) for s in range(steps)])
    want = g2["losses_bn" if bn else "losses"]
    noise = g2["self_noise_bn" if bn else "self_noise"]          # the reference against itself with another BLAS thread count
    per_step = np.abs(got / want - 1).max(axis=1)
    c2["bn" if bn else "plain"] = {"first_step_rel": float(per_step[0]), "max_rel": float(per_step.max()),
                                   "within_4x_reference_self_noise": bool((per_step <= np.maximum(2e-4, 4.0 * noise)).all())}
    del trc, mm

# ---- step time at the run.sh shape, global batch 1024
dims = [2048, 1024, 512, 256, 128, 64]
args2 = make_args(layers=dims, num_emb_list=[256] * 4, e_dim=32, epochs=1)
torch.manual_seed(0)
m2 = RQVAE(in_dim=4096, num_emb_list=[256] * 4, e_dim=32, layers=dims, sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)
x = torch.randn(8 * 1024, 4096, generator=torch.Generator().manual_seed(1))
loader2 = torch.utils.data.DataLoader(x, batch_size=1024, shuffle=False)
tr2 = DataParallelTrainer(args2, m2, len(loader2))
tr2._train_epoch(loader2, 0)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
tr2._train_epoch(loader2, 1)
torch.cuda.synchronize(); dist.barrier()
ms = (time.perf_counter() - t0) / len(loader2) * 1e3
if rank == 0:
    print(json.dumps({"world": world, "max_rel_dev_from_reference_losses": rel,
                      "ok": bool(rel < 2e-3 and replicas_equal and all(v["first_step_rel"] < 2e-5 and v["within_4x_reference_self_noise"] for v in c2.values())),
                      "run_sh_shape_vs_reference_trainer": c2,
                      "collision_rate": coll, "reference_collision_rate": float(g["collision_rate"]), "replicas_equal": replicas_equal,
                      "ms_per_step_global_batch_1024": ms, "items_per_s": 1024 / ms * 1e3}))
dist.barrier()
dist.destroy_process_group()
