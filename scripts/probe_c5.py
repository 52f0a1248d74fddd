"""BASELINE.json configs[4] shape on one GPU: 4 levels x 8192 codes, e_dim 256 (encoder tail 64 -> 256), full generation."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import generate_indices as G, ops
from lcrec_b200.models import RQVAE
from lcrec_b200.synth import synth_items
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
torch.manual_seed(0)
m = RQVAE(in_dim=4096, num_emb_list=[8192] * 4, e_dim=256, layers=[2048, 1024, 512, 256, 128, 64],
          sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50).to(dev).eval()
x = torch.from_numpy(synth_items(n, 4096, n_parents=n // 8, seed=5)).to(dev)
with torch.no_grad():
    z = m.encoder(x[:8192])
    r = z.clone()
    for l, vq in enumerate(m.rq.vq_layers):            # data-driven codebooks: samples of the residual + noise
        idx = torch.randint(0, r.shape[0], (8192,), device=dev)
        vq.embedding.weight.data.copy_(r[idx] + 1e-3 * r.std() * torch.randn(8192, 256, device=dev))
        codes = ops.rq_quantize(r, [vq.embedding.weight.data])["codes"][:, 0]
        r = r - vq.embedding.weight.data[codes]
ix = G.build_indexer(m, n)
for _ in range(2): codes, stats = ix.run_device(x, 20)
torch.cuda.synchronize(); t0 = time.perf_counter()
ops.profile_enable(True); ops.profile_collect()
codes, stats = ix.run_device(x, 20)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
prof = ops.profile_collect(); ops.profile_enable(False)
print(json.dumps({"n": n, "ms": round(dt * 1e3, 2), "items_per_s": round(n / dt), "stats": stats,
                  "stage_ms": {str(k): round(v[0], 3) for k, v in sorted(prof.items())}}))
