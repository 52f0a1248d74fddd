"""Generate tests/golden/pool_kat.npz by running the UNMODIFIED reference embedding producer
(/root/reference/data_process/amazon_text_emb.py generate_item_embedding :49-105, torch CPU) with a stand-in tokenizer
and a stand-in "PLM" (no checkpoint exists offline).  Build container only.

The stand-ins only supply tensors: the tokenizer maps words to ids, pads every sequence to a multiple of 8 (so the
attention mask has zeros) and truncates at max_sent_len; the model returns E[input_ids] + P[position] (+ garbage on
the padded positions, which the mask must remove).  Everything after `outputs = model(...)` - the masked mean pool
(:91-92), the mean over the two text fields (:96), the concatenation and np.save (:100-105) - is the reference's code.
Recorded: the tables, the token ids / masks the reference saw per (item, field), and the saved .npy."""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/data_process"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)
import torch                                   # noqa: E402
import amazon_text_emb as A                    # noqa: E402

H, VOCAB, MAXLEN = 96, 50, 24
LOG = []


class Encoded(dict):
    def to(self, device):
        return self

    @property
    def input_ids(self):
        return self["input_ids"]

    @property
    def attention_mask(self):
        return self["attention_mask"]


def tokenizer(sentences, max_length, truncation, return_tensors, padding):
    rows = [[(sum(map(ord, w)) % (VOCAB - 1)) + 1 for w in s.split(" ")][:max_length] for s in sentences]
    width = (max(len(r) for r in rows) + 7) // 8 * 8
    ids = torch.tensor([r + [0] * (width - len(r)) for r in rows], dtype=torch.int64)
    mask = torch.tensor([[1] * len(r) + [0] * (width - len(r)) for r in rows], dtype=torch.int64)
    return Encoded(input_ids=ids, attention_mask=mask)


def main():
    g = torch.Generator().manual_seed(2024)
    E = torch.randn(VOCAB, H, generator=g)
    P = torch.randn(64, H, generator=g) * 0.1

    def model(input_ids, attention_mask):
        h = E[input_ids] + P[: input_ids.shape[1]][None]
        h = h + (1 - attention_mask).unsqueeze(-1) * 1e3          # padded positions hold garbage
        LOG.append((input_ids.numpy().copy(), attention_mask.numpy().copy()))
        return types.SimpleNamespace(last_hidden_state=h)

    rng = np.random.default_rng(7)
    words = ["w%d" % i for i in range(200)]
    items = []
    for i in range(12):
        title = " ".join(rng.choice(words, size=int(rng.integers(1, 9))))
        desc = " ".join(rng.choice(words, size=int(rng.integers(3, 40))))
        items.append([i, [title, desc]])
    with tempfile.TemporaryDirectory() as tmp:
        args = types.SimpleNamespace(root=tmp, dataset="Toy", plm_name="standin", max_sent_len=MAXLEN, device="cpu")
        A.generate_item_embedding(args, items, tokenizer, model, word_drop_ratio=-1)
        emb = np.load(os.path.join(tmp, "Toy.emb-standin-td.npy"))
    out = {"E": E.numpy(), "P": P.numpy(), "emb": emb, "n_items": np.array(len(items)), "n_fields": np.array(2)}
    for j, (ids, mask) in enumerate(LOG):
        out[f"ids_{j}"] = ids
        out[f"mask_{j}"] = mask
    assert len(LOG) == 2 * len(items)
    np.savez_compressed(os.path.join(OUT, "pool_kat.npz"), **out)
    print("wrote pool_kat.npz", emb.shape, emb.dtype)


if __name__ == "__main__":
    main()
