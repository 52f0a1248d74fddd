"""Multi-GPU hand-over kernels (csrc/exchange.cu) on ONE GPU: four virtual ranks pack their items, the all-to-all is simulated by
copying slabs, the owners unpack, resolve nothing, and send the last codes back.  Checked against a numpy partition: every owner
receives exactly the items of its prefix buckets in ascending global order (bit-exact records), and every origin gets its own
last-level codes back.  (The NCCL path itself is covered by the 2-GPU test and by `sharded_equals_single` of bench.py.)"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200 import _lib
    DEV = torch.device("cuda:0")


def _owner_np(codes, n_codes, world):
    prefix = np.zeros(codes.shape[0], dtype=np.uint64)
    for l in range(codes.shape[1] - 1):
        prefix = prefix * np.uint64(n_codes[l]) + codes[:, l].astype(np.uint64)
    with np.errstate(over="ignore"):
        mixed = ((prefix * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0x7FFFFFFFFFFFFFFF)) >> np.uint64(24)
    return (mixed % np.uint64(world)).astype(np.int64)


@pytest.mark.parametrize("world,sizes,L,D,k", [(4, [5000, 4097, 0, 1], 4, 32, 256), (2, [3000, 3000], 3, 16, 64), (8, [1500] * 8, 4, 256, 8192),
                                               (1, [777], 4, 32, 256)])
def test_exchange_kernels_equal_numpy_partition(world, sizes, L, D, k):
    lib = _lib.load()
    rng = np.random.default_rng(world + L + D)
    n_codes = [k] * L
    p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)
    slab_rows = int(max(sizes) / world * 1.25) + 4096
    slab_bytes = int(lib.lcrec_exchange_slab_bytes(slab_rows, L, D))
    codes, resid, send, slot = [], [], [], []
    for r, n in enumerate(sizes):
        c = rng.integers(0, k, size=(n, L)).astype(np.int64)
        if n > 10:
            c[:, :L - 1] = c[rng.integers(0, max(n // 6, 1), size=n), :L - 1]       # shared prefixes: real buckets
        x = rng.standard_normal((n, D)).astype(np.float32)
        codes.append(c); resid.append(x)
        cd, xd = torch.from_numpy(c).to(DEV), torch.from_numpy(x).to(DEV)
        sb = torch.zeros(slab_bytes * world, dtype=torch.uint8, device=DEV)
        sl = torch.full((max(n, 1),), -7, dtype=torch.int32, device=DEV)
        ws = torch.empty(int(lib.lcrec_exchange_workspace_bytes(max(n, 1), world)), dtype=torch.uint8, device=DEV)
        _lib.check(lib.lcrec_exchange_pack(p(cd) if n else None, p(xd) if n else None, n, L, D, _lib.i32_array(n_codes), world, slab_rows,
                                           p(sb), p(sl), None, p(ws), ws.numel(), st))
        send.append(sb); slot.append(sl)
    all_codes = np.concatenate(codes); all_resid = np.concatenate(resid)
    owner = _owner_np(all_codes, n_codes, world)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    subs = []
    for r in range(world):                                           # the all-to-all: rank r receives slab r of every source
        recv = torch.cat([send[s].view(world, slab_bytes)[r] for s in range(world)])
        hdr = recv.view(world, slab_bytes)[:, :16].contiguous().view(torch.int64).cpu().numpy()
        assert (hdr[:, 1] == 0).all()
        want_idx = np.nonzero(owner == r)[0]                         # ascending global order
        assert int(hdr[:, 0].sum()) == len(want_idx)
        for s in range(world):
            assert int(hdr[s, 0]) == int(((want_idx >= starts[s]) & (want_idx < starts[s + 1])).sum())
        n_mine = len(want_idx)
        sc = torch.full((max(n_mine, 1), L), -1, dtype=torch.int64, device=DEV)
        sr = torch.zeros((max(n_mine, 1), D), dtype=torch.float32, device=DEV)
        _lib.check(lib.lcrec_exchange_unpack(p(recv), world, slab_rows, L, D, p(sc), p(sr), n_mine, st))
        assert np.array_equal(sc.cpu().numpy()[:n_mine], all_codes[want_idx])
        assert np.array_equal(sr.cpu().numpy()[:n_mine], all_resid[want_idx])
        subs.append((recv, sc, n_mine, want_idx))
    # way back: owners change the last code, origins must see exactly those values at their own rows
    new_last = (all_codes[:, -1] * 7 + 3) % k
    back = []
    for r in range(world):
        recv, sc, n_mine, want_idx = subs[r]
        if n_mine:
            sc[:n_mine, -1] = torch.from_numpy(new_last[want_idx]).to(DEV)
        b = torch.full((world, slab_rows), -5, dtype=torch.int64, device=DEV)
        _lib.check(lib.lcrec_exchange_pack_last(p(sc), L, p(recv), world, slab_rows, D, p(b), n_mine, st))
        back.append(b)
    for s, n in enumerate(sizes):
        brecv = torch.stack([back[r][s] for r in range(world)]).contiguous()         # origin s receives slab s of every owner
        out = torch.from_numpy(codes[s]).to(DEV).clone() if n else torch.zeros((1, L), dtype=torch.int64, device=DEV)
        _lib.check(lib.lcrec_exchange_scatter_last(p(brecv), p(slot[s]), n, L, p(out), st))
        if n:
            got = out.cpu().numpy()
            assert np.array_equal(got[:, :-1], codes[s][:, :-1])
            assert np.array_equal(got[:, -1], new_last[starts[s]:starts[s + 1]])


def test_exchange_overflow_is_flagged():
    lib = _lib.load()
    p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)
    n, L, D, world, slab_rows = 5000, 4, 32, 2, 100                 # far too small: both slabs overflow
    c = torch.randint(0, 256, (n, L), device=DEV)
    x = torch.randn(n, D, device=DEV)
    slab_bytes = int(lib.lcrec_exchange_slab_bytes(slab_rows, L, D))
    sb = torch.zeros(slab_bytes * world, dtype=torch.uint8, device=DEV)
    sl = torch.empty(n, dtype=torch.int32, device=DEV)
    ws = torch.empty(int(lib.lcrec_exchange_workspace_bytes(n, world)), dtype=torch.uint8, device=DEV)
    _lib.check(lib.lcrec_exchange_pack(p(c), p(x), n, L, D, _lib.i32_array([256] * L), world, slab_rows, p(sb), p(sl), None, p(ws), ws.numel(), st))
    hdr = sb.view(world, slab_bytes)[:, :16].contiguous().view(torch.int64).cpu().numpy()
    assert (hdr[:, 1] == 1).all() and int(hdr[:, 0].sum()) == n
    assert int((sl < 0).sum()) == n - 2 * slab_rows                  # exactly the rows that did not fit are marked
