"""B200 parity of the embedding producer hand-off (reference data_process/amazon_text_emb.py:49-105):
masked mean pool + field mean against the oracle and the .npy written by the unmodified reference producer
(tests/golden/pool_kat.npz).  Floating point (summation order over the sequence differs between torch's own back ends):
bar 1e-6 of the row scale, far inside the 1e-5 relative bar of the task."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import lcrec_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200 import ops
    from lcrec_b200.text_emb import generate_item_embedding
    DEV = torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


class _Encoded(dict):
    def to(self, device):
        return _Encoded({k: v.to(device) for k, v in self.items()})
    input_ids = property(lambda self: self["input_ids"])
    attention_mask = property(lambda self: self["attention_mask"])


@pytest.mark.parametrize("batch_size", [1])
def test_producer_matches_reference_npy(golden, tmp_path, batch_size):
    """generate_item_embedding with stand-ins that replay the token ids the reference run saw: same file name, values
    within 1e-6 of the row scale, and the returned device matrix is what was saved."""
    g = golden("pool_kat")
    n, nf = int(g["n_items"]), int(g["n_fields"])
    calls = iter(range(n * nf))
    E, P = T(g["E"]), T(g["P"])

    def tokenizer(sentences, max_length, truncation, return_tensors, padding):
        j = next(calls)
        assert max_length == 24 and truncation and return_tensors == "pt" and padding == "longest"
        return _Encoded(input_ids=torch.from_numpy(g[f"ids_{j}"]), attention_mask=torch.from_numpy(g[f"mask_{j}"]))

    def model(input_ids, attention_mask):
        h = E[input_ids] + P[: input_ids.shape[1]][None] + (1 - attention_mask).unsqueeze(-1) * 1e3
        return types.SimpleNamespace(last_hidden_state=h)

    items = [[i, ["title %d" % i, "description %d" % i]] for i in reversed(range(n))]      # order_texts re-sorts by id
    args = types.SimpleNamespace(root=str(tmp_path), dataset="Toy", plm_name="standin", max_sent_len=24, device=DEV)
    emb = generate_item_embedding(args, items, tokenizer, model, word_drop_ratio=-1, batch_size=batch_size)
    saved = np.load(os.path.join(str(tmp_path), "Toy.emb-standin-td.npy"))
    assert emb.dtype == torch.float32 and emb.is_cuda and saved.dtype == np.float32
    assert np.array_equal(saved, emb.cpu().numpy()) and saved.shape == g["emb"].shape
    scale = np.abs(g["emb"]).max(axis=1, keepdims=True)
    assert (np.abs(saved - g["emb"]) <= 1e-6 * scale).all()


@pytest.mark.parametrize("b,t,h,dtype", [(1, 2048, 4096, "float32"), (3, 37, 100, "float32"), (64, 128, 768, "float32"),
                                         (2, 5, 8, "float32"), (1, 1, 4096, "float32"), (5, 300, 1024, "float16"),
                                         (5, 300, 1024, "bfloat16"), (4, 33, 250, "float16"), (300, 16, 64, "float32")])
def test_masked_mean_pool_matches_oracle(b, t, h, dtype):
    """Sequence splits (small batch x long sequence), the scalar path (hidden size not a multiple of the vector width),
    ragged masks incl. full rows and single tokens, half-precision inputs (accumulated in fp32 like the oracle on the
    up-cast input)."""
    rng = np.random.default_rng(b * 1000 + t + h)
    x = torch.from_numpy(rng.standard_normal((b, t, h)).astype(np.float32)).to(getattr(torch, dtype))
    lens = rng.integers(1, t + 1, size=b)
    lens[0] = t
    mask = (np.arange(t)[None, :] < lens[:, None]).astype(np.int64)
    x = torch.where(torch.from_numpy(mask).bool()[..., None], x, torch.full_like(x, 1e3))     # garbage on padding
    want = O.masked_mean_pool(x.float().numpy(), mask)
    got = ops.masked_mean_pool(x.to(DEV), T(mask)).cpu().numpy()
    scale = np.abs(want).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(got - want) <= 2e-6 * scale).all(), np.abs(got - want).max()
    if t <= 16:
        assert np.array_equal(got, want)                       # one split, position order: bit-identical


def test_masked_mean_pool_field_mean_into_matrix_rows():
    """accumulate / divide_by and writing into a row slice of a wider, strided embedding matrix."""
    rng = np.random.default_rng(5)
    hs = [rng.standard_normal((6, 20, 128)).astype(np.float32) for _ in range(3)]
    ms = [(np.arange(20)[None] < rng.integers(1, 21, size=6)[:, None]).astype(np.int64) for _ in range(3)]
    want = O.item_embedding(hs, ms)
    big = torch.full((10, 256), -7.0, device=DEV)
    rows = big[2:8, 64:192]
    for f in range(3):
        ops.masked_mean_pool(T(hs[f]), T(ms[f]), out=rows, accumulate=f > 0, divide_by=3.0 if f == 2 else 0.0)
    got = rows.cpu().numpy()
    assert (np.abs(got - want) <= 2e-6 * np.abs(want).max(axis=1, keepdims=True)).all()
    untouched = big.clone()
    untouched[2:8, 64:192] = -7.0
    assert (untouched == -7.0).all()
    with pytest.raises(RuntimeError):
        ops.masked_mean_pool(T(hs[0]).double(), T(ms[0]))
