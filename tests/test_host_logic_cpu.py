"""CPU checks of the host-side mirror of the reference interface (no kernels run)."""
import argparse
import json

import numpy as np
import pytest
import torch

from lcrec_b200 import generate_indices as G
from lcrec_b200.datasets import EmbDataset
from lcrec_b200.main import parse_args
from lcrec_b200.models import RQVAE, MLPLayers, VectorQuantizer, ResidualVectorQuantizer
from lcrec_b200.models.layers import activation_layer
from lcrec_b200.trainer import _constant_warmup, _linear_warmup_decay
from oracle import lcrec_oracle as O
from tests.conftest import state_dict_of


def small(bn=False):
    return RQVAE(in_dim=96, num_emb_list=[32] * 4, e_dim=16, layers=[64, 48], bn=bn,
                 sk_epsilons=[0.0, 0.0, 0.0, 0.003], sk_iters=50)


@pytest.mark.parametrize("name,bn", [("small_model", False), ("bn_model", True)])
def test_reference_checkpoints_load(golden, name, bn):
    g = golden(name)
    sd = {k: torch.from_numpy(v) for k, v in state_dict_of(g).items()}
    m = small(bn)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)


def test_constructor_surface():
    vq = VectorQuantizer(32, 16, beta=0.3, kmeans_init=True, kmeans_iters=7, sk_epsilon=0.01, sk_iters=9)
    for attr in ("n_e", "e_dim", "beta", "kmeans_init", "kmeans_iters", "sk_epsilon", "sk_iters", "initted", "embedding"):
        assert hasattr(vq, attr)
    assert vq.initted is False and float(vq.embedding.weight.abs().sum()) == 0.0
    vq2 = VectorQuantizer(32, 16)
    assert vq2.initted and float(vq2.embedding.weight.abs().max()) <= 1 / 32
    assert vq2.get_codebook() is vq2.embedding.weight
    assert vq2.get_codebook_entry(torch.tensor([1, 2]), shape=(2, 16)).shape == (2, 16)
    rq = ResidualVectorQuantizer([8, 8], 4, sk_epsilons=[0.0, 0.003])
    assert rq.get_codebook().shape == (2, 8, 4) and rq.vq_layers[1].sk_epsilon == 0.003
    m = small()
    assert m.encode_layer_dims == [96, 64, 48, 16] and m.decode_layer_dims == [16, 48, 64, 96]
    with pytest.raises(ValueError, match="incompatible loss type"):
        RQVAE(in_dim=8, num_emb_list=[4], e_dim=4, layers=[4], sk_epsilons=[0.0], loss_type="huber").compute_loss(
            torch.zeros(1), torch.zeros(()), xs=torch.zeros(1))


def test_activation_factory():
    assert isinstance(activation_layer("ReLU"), torch.nn.ReLU)
    assert activation_layer("none") is None and activation_layer(None) is None
    assert isinstance(activation_layer(torch.nn.Tanh), torch.nn.Tanh)
    with pytest.raises(NotImplementedError):
        activation_layer(3)


def test_bn_folding_matches_eval_forward(golden):
    g = golden("bn_model")
    m = small(True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state_dict_of(g).items()})
    m.eval()
    ws, bs = m.encoder._folded()
    x = torch.from_numpy(g["x"][:64])
    h = x
    for i, (w, b) in enumerate(zip(ws, bs)):
        h = h @ w.t() + b
        if i != len(ws) - 1:
            h = torch.relu(h)
    with torch.no_grad():
        ref = m.encoder.mlp_layers(x)
    np.testing.assert_allclose(h.numpy(), ref.numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(h.numpy(), g["latents"][:64], rtol=2e-5, atol=2e-6)


def test_generation_epsilon_rule():
    m = small()
    m.rq.vq_layers[3].sk_epsilon = 0.0
    assert G.apply_generation_epsilons(m) == 0.003
    assert [v.sk_epsilon for v in m.rq.vq_layers] == [0.0, 0.0, 0.0, 0.003]
    m.rq.vq_layers[3].sk_epsilon = 0.01
    assert G.apply_generation_epsilons(m) == 0.01


def test_index_json_is_byte_identical_to_reference(golden, tmp_path):
    g = golden("small_model")
    out = tmp_path / "x.index.json"
    G.write_index_json(g["script_codes_final"], str(out))
    assert out.read_bytes() == g["script_json"].tobytes()
    assert json.loads(out.read_text())["0"][0].startswith("<a_")


def test_index_json_satisfies_the_downstream_consumer(golden, tmp_path):
    """The file is read by BaseDataset (reference data.py:38-89): json.load -> {str(item): [tok_0..tok_{L-1}]};
    get_new_tokens = sorted set of all tokens (fed to tokenizer.add_tokens), get_all_items = set of "".join(tokens)
    (must be one entry per item when the table is collision free), get_prefix_allowed_tokens_fn groups tokens by
    level.  Restated here on our writer's output."""
    import re
    g = golden("small_model")
    codes = np.asarray(g["script_codes_final"])
    out = tmp_path / "y.index.json"
    G.write_index_json(codes, str(out))
    indices = json.loads(out.read_text())
    assert list(indices.keys()) == [str(i) for i in range(codes.shape[0])]            # int keys serialised as strings, item order
    L = codes.shape[1]
    pat = [re.compile(r"^<%s_(\d+)>$" % "abcde"[l]) for l in range(L)]
    new_tokens, all_items, by_level = set(), set(), {}
    for idx, (item, toks) in enumerate(indices.items()):
        assert len(toks) == L
        for l, t in enumerate(toks):
            m = pat[l].match(t)
            assert m and int(m.group(1)) == int(codes[idx, l])                          # level letter + code value round-trip
            new_tokens.add(t)
            by_level.setdefault(l, set()).add(t)
        all_items.add("".join(toks))
    assert sorted(new_tokens) == sorted(set(t for toks in indices.values() for t in toks))
    assert len(all_items) == len({tuple(r) for r in codes.tolist()})                   # one string per distinct code tuple
    assert all(len(by_level[l]) == len(set(codes[:, l].tolist())) for l in range(L))  # allowed-token sets per level
    assert not (set.intersection(*[by_level[l] for l in range(L)]) if L > 1 else set())   # levels never share a token


def test_schedulers_match_transformers():
    tr = pytest.importorskip("transformers")
    for make_ours, make_ref in [
        (lambda o: _linear_warmup_decay(o, 5, 40), lambda o: tr.get_linear_schedule_with_warmup(o, 5, 40)),
        (lambda o: _constant_warmup(o, 5), lambda o: tr.get_constant_schedule_with_warmup(o, 5)),
    ]:
        lrs = []
        for make in (make_ours, make_ref):
            p = torch.nn.Parameter(torch.zeros(1))
            opt = torch.optim.SGD([p], lr=1.0)
            s = make(opt)
            seq = []
            for _ in range(45):
                seq.append(opt.param_groups[0]["lr"]); opt.step(); s.step()
            lrs.append(seq)
        assert lrs[0] == lrs[1]


def test_cli_bool_quirk():
    a = parse_args(["--bn", "False", "--sk_epsilons", "0", "0", "0", "0.003", "--num_emb_list", "256", "256", "256", "256"])
    assert a.bn is True and a.kmeans_init is True and a.sk_iters == 50 and a.batch_size == 2048


def test_embdataset_list_index(tmp_path):
    x = np.arange(40, dtype=np.float32).reshape(10, 4)
    np.save(tmp_path / "e.npy", x)
    d = EmbDataset(str(tmp_path / "e.npy"))
    assert d.dim == 4 and len(d) == 10
    assert d[3].dtype == torch.float32 and d[[1, 5]].shape == (2, 4)
    assert torch.equal(d[[1, 5]], torch.from_numpy(x[[1, 5]]))
