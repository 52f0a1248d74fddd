"""world_size-2 gloo test of the data-parallel trainer's host logic (lcrec_b200.dp_trainer) on CPU: row sharding of every
global batch, loss weighting, the single flat gradient all-reduce, identical clipping + optimiser step on every rank.
The model is a plain-torch stand-in with the RQVAE call surface (the real model has no CPU path); what is under test is
that two ranks reproduce the single-process step on the global batch."""
import os
import socket
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class StandIn(torch.nn.Module):
    """forward(x) -> (out, quant_loss, indices); compute_loss(out, quant_loss, xs) -> (total, recon): the RQVAE surface."""
    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Linear(12, 5)
        self.dec = torch.nn.Linear(5, 12)
        self.codebook = torch.nn.Parameter(torch.randn(7, 5) * 0.3)
        self.unused = torch.nn.Parameter(torch.zeros(3))           # never receives a gradient
        self.rq = types.SimpleNamespace(vq_layers=[types.SimpleNamespace(initted=False, sk_epsilon=0.0)])
        self.init_calls = []                                        # (rows, use_sk) of the k-means initialisation passes

    def forward(self, x, use_sk=True):
        if not self.rq.vq_layers[0].initted:                        # vq.py:67-68: the first batch initialises the codebook
            self.init_calls.append((x.shape[0], use_sk, torch.is_grad_enabled()))
            self.rq.vq_layers[0].initted = True
        z = self.enc(x)
        idx = torch.cdist(z.detach(), self.codebook.detach()).argmin(dim=1)
        q = self.codebook[idx]
        qloss = torch.nn.functional.mse_loss(q, z.detach()) + 0.25 * torch.nn.functional.mse_loss(q.detach(), z)
        return self.dec(z + (q - z).detach()), qloss, idx

    def compute_loss(self, out, quant_loss, xs=None):
        recon = torch.nn.functional.mse_loss(out, xs)
        return recon + quant_loss, recon


def _args(tmp):
    return types.SimpleNamespace(lr=1e-2, learner="AdamW", lr_scheduler_type="linear", weight_decay=1e-4, epochs=2,
                                 warmup_epochs=1, save_limit=2, eval_step=2, device="cpu", ckpt_dir=tmp)


def _batches():
    g = torch.Generator().manual_seed(5)
    return [torch.randn(n, 12, generator=g) for n in (16, 9, 1, 16)]   # ragged: 9 -> 5 + 4 rows, 1 -> 1 + 0 rows


def _worker(rank, world, port, tmp, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lcrec_b200.dp_trainer import DataParallelTrainer
    torch.manual_seed(100 + rank)                                  # different local init: rank 0's weights are broadcast
    model = StandIn()
    tr = DataParallelTrainer(_args(tmp), model, data_num=4)
    losses = [tr._train_epoch(_batches(), ep) for ep in range(2)]
    path = tr._save_checkpoint(0, ckpt_file="dp.pth")
    dist.barrier()
    out[rank] = (losses, [p.detach().clone() for p in model.parameters()], os.path.exists(path), model.init_calls)
    dist.destroy_process_group()


def test_two_rank_training_equals_single_process(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, str(tmp_path), out), nprocs=2, join=True)
    # single process on the global batches, same initial weights (seed of rank 0)
    from lcrec_b200.trainer import Trainer
    torch.manual_seed(100)
    model = StandIn()
    tr = Trainer(_args(str(tmp_path)), model, data_num=4)
    ref_losses = [tr._train_epoch(_batches(), ep) for ep in range(2)]
    for rank in (0, 1):
        losses, params, saved, init_calls = out[rank]
        assert init_calls == [(16, True, False)]                   # ONE pass over the full first global batch, the step's use_sk, no grad
        np.testing.assert_allclose(np.array(losses), np.array(ref_losses), rtol=2e-5)
        for a, b in zip(params, model.parameters()):
            np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=2e-4, atol=2e-6)
        assert saved                                               # rank 0 wrote it; every rank sees the path
    for a, b in zip(out[0][1], out[1][1]):
        assert torch.equal(a, b)                                   # replicas stay bit-identical


def test_dp_trainer_needs_a_process_group(tmp_path):
    import pytest
    from lcrec_b200.dp_trainer import DataParallelTrainer
    with pytest.raises(RuntimeError, match="process group"):
        DataParallelTrainer(_args(str(tmp_path)), StandIn(), 4)
