"""`.index.json` emitter on the device (lcrec_index_json) == json.dump of the reference's dict (generate_indices.py:138-145)."""
import json

import numpy as np
import pytest
import torch

from oracle import lcrec_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200 import ops
    from lcrec_b200 import generate_indices as G
    DEV = torch.device("cuda:0")


@pytest.mark.parametrize("n,L,k", [(0, 4, 256), (1, 4, 256), (2, 1, 10), (1023, 4, 256), (1024, 4, 256), (1025, 5, 8192),
                                   (12345, 3, 256), (300000, 4, 256), (1100000, 4, 8192)])
def test_device_json_is_byte_identical(n, L, k, tmp_path):
    rng = np.random.default_rng(n + L)
    codes = rng.integers(0, k, size=(n, L)).astype(np.int64)
    if n > 10:
        codes[3] = 0; codes[4] = k - 1                   # shortest / longest tokens
    got = ops.index_json_bytes(torch.from_numpy(codes).to(DEV))
    want = json.dumps({i: [G.PREFIX[l].format(int(v)) for l, v in enumerate(row)] for i, row in enumerate(codes.tolist())}).encode()
    assert len(got) == len(want)
    assert got == want
    if n and n < 20000:
        assert got.decode() == O.index_json(codes)
    G.write_index_json(torch.from_numpy(codes).to(DEV), str(tmp_path / "a.json"))
    assert (tmp_path / "a.json").read_bytes() == want
