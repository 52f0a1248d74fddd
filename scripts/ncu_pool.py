"""Three launches of the masked mean pool (64 x 512 x 4096 fp32, 537 MB) for an `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcrec_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(64, 512, 4096, device=dev)
mask = torch.ones(64, 512, dtype=torch.int64, device=dev)
out = torch.empty(64, 4096, device=dev)
for _ in range(3):
    ops.masked_mean_pool(x, mask, out=out)
torch.cuda.synchronize()
print("ok")
