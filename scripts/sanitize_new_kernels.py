"""One small invocation of every kernel added for the SURVEY 8(f) rows and the training path (odd sizes, empty tensors, list
flushes, sequence splits, relocation) - a quick crash test on a GPU box.  compute-sanitizer is closed on this pool
(`gpurun_out/sanitize_new.log`), so bounds are covered by the parity tests at ragged shapes instead."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcrec_b200 import ops
from lcrec_b200.optim import FusedAdam
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)

# EMA: small, and one code with more members than the list buffer (several flushes), e_dim not a multiple of 32
for n, k, d, skew in ((300, 16, 8, 0.0), (5000, 4, 40, 0.9)):
    idx = torch.randint(0, k, (n,), device=dev, generator=g)
    idx[torch.rand(n, device=dev, generator=g) < skew] = 1
    cs, w, cb = torch.zeros(k, device=dev), torch.zeros(k, d, device=dev), R(k, d)
    ops.ema_update(R(n, d), idx, cs, w, cb, 0.99, 1e-5)
    ops.codebook_usage(cs, 1e-5, 1e-5, want_unused=True)
# pool: vector / scalar / half paths, sequence splits, strided output rows
ops.masked_mean_pool(R(2, 40, 64), torch.ones(2, 40, dtype=torch.int64, device=dev))
ops.masked_mean_pool(R(3, 37, 50), (torch.arange(37, device=dev)[None] < torch.tensor([[37], [1], [20]], device=dev)).long())
ops.masked_mean_pool(R(1, 300, 72).half(), torch.ones(1, 300, dtype=torch.int64, device=dev))
big = torch.zeros(5, 200, device=dev)
ops.masked_mean_pool(R(3, 9, 64).bfloat16(), torch.ones(3, 9, dtype=torch.int64, device=dev), out=big[1:4, 10:74])
# k-means incl. the empty-cluster relocation
x = R(600, 20) + 2.0
xc, mean, var = ops.kmeans_center(x)
seeds = xc[:10].clone(); seeds[3] += 50.0
ops.kmeans_lloyd(xc, seeds, 20, 1e-4 * var, add_mean=mean)
# training-side quantiser
cbs = [R(9, 48), R(5, 48)]
codes = torch.stack([torch.randint(0, 9, (257,), device=dev, generator=g), torch.randint(0, 5, (257,), device=dev, generator=g)], 1)
r = ops.rq_train_forward(R(257, 48), codes, cbs)
ops.rq_train_backward(r["diffs"], r["codes_t"], [9, 5], R(257, 48), torch.ones((), device=dev), 0.25)
# optimiser: more tensors than one argument table, odd sizes, an empty tensor
ps = [torch.nn.Parameter(R(*s)) for s in [(4097,), (33, 7), (0,), (1,)] + [(5, 3)] * 40]
for p in ps:
    p.grad = torch.randn_like(p)
opt = FusedAdam(ps, lr=1e-3, weight_decay=1e-4)
opt.clip_and_step(1.0, want_norm=True)
opt.step()
torch.cuda.synchronize()
print("sanitize script done")
