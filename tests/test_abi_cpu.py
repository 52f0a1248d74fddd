"""CPU checks of the boundary: the C-ABI library loads, exports every symbol the header declares,
and refuses to compute without an sm_100 device (no CPU fallback anywhere)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lcrec_b200.h")


@pytest.fixture(scope="module")
def lib():
    from lcrec_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lcrec_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lcrec_b200.h but not exported"


def test_ctypes_signatures_cover_header(lib):
    from lcrec_b200 import _lib
    assert set(header_functions()) - set(_lib.SIGNATURES) == set()


def test_version_and_strerror(lib):
    assert lib.lcrec_version() >= 100
    assert lib.lcrec_strerror(0) == b"ok"
    assert b"argument" in lib.lcrec_strerror(1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(lib):
    rc = lib.lcrec_device_check()
    assert rc != 0
    assert len(lib.lcrec_last_error()) > 0
    from lcrec_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.linear_forward(torch.zeros(4, 8), torch.zeros(4, 8), None, True)
    with pytest.raises(RuntimeError):
        ops.rq_quantize(torch.zeros(4, 8), [torch.zeros(4, 8)])
    from lcrec_b200.models import MLPLayers
    with pytest.raises(RuntimeError, match="CUDA only"):
        MLPLayers([8, 4, 4])(torch.zeros(2, 8))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lcrec_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lcrec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "lcrec_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_argument_validation_of_late_entry_points(lib):
    """Bad arguments are refused with LCREC_ERR_ARG (1) before anything touches the device, with a message that names
    the violated condition; valid arguments on a machine without a B200 fail with a CUDA / unsupported-device code."""
    import ctypes as C
    null = C.c_void_p(0)
    one = C.c_void_p(8)                       # a non-null dummy pointer: never dereferenced on these paths
    cases = {
        "ema n < 0": lambda: lib.lcrec_ema_update(one, one, -1, 4, 8, 0.99, 1e-5, one, one, one, null),
        "ema e_dim too large": lambda: lib.lcrec_ema_update(one, one, 4, 4, 9000, 0.99, 1e-5, one, one, one, null),
        "ema null buffers": lambda: lib.lcrec_ema_update(one, one, 4, 4, 8, 0.99, 1e-5, null, one, one, null),
        "usage null": lambda: lib.lcrec_codebook_usage(null, 4, 1e-5, 1e-5, one, null, null),
        "pool dtype": lambda: lib.lcrec_masked_mean_pool(one, 7, one, 2, 3, 8, one, 8, 0, 0.0, null, 0, null),
        "pool stride": lambda: lib.lcrec_masked_mean_pool(one, 0, one, 2, 3, 8, one, 4, 0, 0.0, null, 0, null),
        "pool seq_len 0": lambda: lib.lcrec_masked_mean_pool(one, 0, one, 2, 0, 8, one, 8, 0, 0.0, null, 0, null),
        "kmeans center n 0": lambda: lib.lcrec_kmeans_center(one, 0, 8, one, one, None, null, 0, null),
        "kmeans lloyd max_iter 0": lambda: lib.lcrec_kmeans_lloyd(one, 4, 8, one, 2, 0, 1e-4, null, null, None, None, null, 0, null),
        "kmeanspp first index": lambda: lib.lcrec_kmeanspp_seed(one, 4, 8, 2, 9, one, 3, one, null, null, 0, null),
        "kmeanspp trials": lambda: lib.lcrec_kmeanspp_seed(one, 4, 8, 2, 0, one, 99, one, null, null, 0, null),
        "rq train levels": lambda: lib.lcrec_rq_train_forward(one, one, 4, 8, 9, None, one, one, one, one, null),
        "rq train backward n": lambda: lib.lcrec_rq_train_backward(one, one, 0, 8, 1, (C.c_int32 * 1)(4), null, null, 0.25, null, None, null),
        "adam step 0": lambda: lib.lcrec_adam_clip_step(1, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 0, 1.0, 1, null, null, 0, null),
    }
    for name, call in cases.items():
        assert call() == 1, name
        assert b"bad argument" in lib.lcrec_last_error(), name
    assert lib.lcrec_masked_mean_pool(one, 0, one, 0, 3, 8, one, 8, 0, 0.0, null, 0, null) == 0      # empty batch: nothing to do
    assert lib.lcrec_adam_clip_step(0, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1, 1.0, 1, null, null, 0, null) == 0
    if not torch.cuda.is_available():
        rc = lib.lcrec_ema_update(one, one, 4, 4, 8, 0.99, 1e-5, one, one, one, null)
        assert rc in (2, 3) and len(lib.lcrec_last_error()) > 0                                        # no device: loud failure
    assert lib.lcrec_kmeans_workspace_bytes(1024, 32, 256) > 3 * 1024 * 8
    assert lib.lcrec_masked_mean_pool_workspace_bytes(1, 2048, 4096) >= 128 * 4096 * 4
    assert lib.lcrec_kmeanspp_workspace_bytes(1024, 7) > 7 * 1024 * 4
