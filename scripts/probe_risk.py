import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrec_b200 import ops, _lib
import ctypes as C
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(2)
ng = 20000
sizes = rng.integers(2, 3, size=ng)
n_items = int(sizes.sum())
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
rng2 = np.random.default_rng(2)
centres = (rng2.standard_normal((ng, 32)) * 0.05).astype(np.float32)
resid = np.repeat(centres, sizes, axis=0) + (rng2.standard_normal((n_items, 32)) * 0.003).astype(np.float32)
cb = (rng2.standard_normal((256, 32)) * 0.05).astype(np.float32)
mem = np.arange(n_items, dtype=np.int64)
lib = _lib.load()
ops.sinkhorn_set_mode(2)
codes = torch.zeros((n_items, 4), dtype=torch.int64, device=dev)
r, cbt = T(resid), T(cb)
flags = torch.zeros(2, dtype=torch.int32, device=dev)
nbytes = lib.lcrec_sinkhorn_groups_workspace_bytes(n_items, 256)
ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
offt, memt, ngt = T(off), T(mem), torch.tensor([ng], device=dev)
_lib.check(lib.lcrec_sinkhorn_groups_part(C.c_void_p(r.data_ptr()), 32, C.c_void_p(cbt.data_ptr()), 256, C.c_void_p(offt.data_ptr()),
            C.c_void_p(memt.data_ptr()), C.c_void_p(ngt.data_ptr()), ng, n_items, 0.003, 50,
            C.c_void_p(codes.data_ptr()), 4, 3, 1, 0, C.c_void_p(flags.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), None))
torch.cuda.synchronize()
cnt = int(ws[8:12].view(torch.int32).item())
cap = min(n_items, 1 << 20)
off_big = 256
off_risky = ((off_big + 8 * cap * 257 + 255) // 256) * 256
lst = ws[off_risky: off_risky + 4 * cnt].view(torch.int32).cpu().numpy()
reasons = (lst >> 27) & 15
print("risky", cnt, "of", ng, "reasons histogram (bit1 band<one, bit2 noone_band, bit4 noone_near, bit8 nan):", np.bincount(reasons, minlength=16).tolist())
print("first groups", (lst[:10] & 0x07FFFFFF).tolist())
