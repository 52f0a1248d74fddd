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
