"""Training loop with the interface of reference ``index/trainer.py`` (Trainer :14-251).

Same constructor (``Trainer(args, model, data_num)``), optimiser / scheduler choices, gradient
clipping, NaN guard, summed-loss bookkeeping, checkpoint dictionary and rotation.  The model's
forward runs on the lcrec_b200 kernels; the collision-rate evaluation (:127-152) uses the device
sort/unique (``lcrec_collisions``) instead of a Python set of strings.
"""
from __future__ import annotations

import heapq
import logging
import os
from time import time

import numpy as np
import torch
from torch import optim
from torch.optim.lr_scheduler import LambdaLR
from tqdm import tqdm

from . import ops
from .optim import FusedAdam
from .utils import delete_file, ensure_dir, get_local_time, set_color


# Adam / AdamW through lcrec_adam_clip_step (LCREC_FUSED_OPTIM=0: torch.optim + clip_grad_norm_)
FUSED_OPTIM = os.environ.get("LCREC_FUSED_OPTIM", "1") != "0"
# the whole step captured in a CUDA graph per batch size (LCREC_TRAIN_GRAPH=0: eager launches, as in round 1)
TRAIN_GRAPH = os.environ.get("LCREC_TRAIN_GRAPH", "1") != "0"


def _linear_warmup_decay(optimizer, warmup, total):
    """Same multiplier as transformers.get_linear_schedule_with_warmup (trainer.py:84-87)."""
    def f(step):
        if step < warmup:
            return float(step) / float(max(1, warmup))
        return max(0.0, float(total - step) / float(max(1, total - warmup)))
    return LambdaLR(optimizer, f)


def _constant_warmup(optimizer, warmup):
    """Same multiplier as transformers.get_constant_schedule_with_warmup (trainer.py:88-90)."""
    def f(step):
        return float(step) / float(max(1.0, warmup)) if step < warmup else 1.0
    return LambdaLR(optimizer, f)


class Trainer(object):
    def __init__(self, args, model, data_num):
        self.args = args
        self.model = model
        self.logger = logging.getLogger()
        self.lr = args.lr
        self.learner = args.learner
        self.lr_scheduler_type = args.lr_scheduler_type
        self.weight_decay = args.weight_decay
        self.epochs = args.epochs
        self.warmup_steps = args.warmup_epochs * data_num
        self.max_steps = args.epochs * data_num
        self.save_limit = args.save_limit
        self.best_save_heap = []
        self.newest_save_queue = []
        self.eval_step = min(args.eval_step, self.epochs)
        self.device = torch.device(args.device)
        self.ckpt_dir = os.path.join(args.ckpt_dir, "{}".format(get_local_time()))
        ensure_dir(self.ckpt_dir)
        self.best_loss = np.inf
        self.best_collision_rate = np.inf
        self.best_loss_ckpt = "best_loss_model.pth"
        self.best_collision_ckpt = "best_collision_model.pth"
        self.optimizer = self._build_optimizer()
        self.scheduler = self._get_scheduler()
        self.model = self.model.to(self.device)

    def _build_optimizer(self):
        params = self.model.parameters()
        name = self.learner.lower()
        kw = dict(lr=self.lr, weight_decay=self.weight_decay)
        table = {"adam": optim.Adam, "sgd": optim.SGD, "adagrad": optim.Adagrad, "rmsprop": optim.RMSprop,
                 "adamw": optim.AdamW}
        if name in ("adam", "adamw") and FUSED_OPTIM and self.device.type == "cuda":
            # same update rule and state layout as torch.optim.Adam / AdamW; clipping + step in two launches
            return FusedAdam(params, decoupled=(name == "adamw"), **kw)
        if name in ("sgd", "adagrad", "rmsprop") and FUSED_OPTIM and self.device.type == "cuda":
            from .optim import FusedSimple                     # trainer.py:62-75 with clip_grad_norm_ fused in
            return FusedSimple(params, name, **kw)
        if name in table:
            opt = table[name](params, **kw)
            if name == "adagrad":
                for state in opt.state.values():
                    for k, v in state.items():
                        if torch.is_tensor(v):
                            state[k] = v.to(self.device)
            return opt
        self.logger.warning("Received unrecognized optimizer, set default Adam optimizer")
        return optim.Adam(params, lr=self.lr)

    def _get_scheduler(self):
        if self.lr_scheduler_type.lower() == "linear":
            return _linear_warmup_decay(self.optimizer, self.warmup_steps, self.max_steps)
        return _constant_warmup(self.optimizer, self.warmup_steps)

    def _check_nan(self, loss):
        if torch.isnan(loss):
            raise ValueError("Training loss is nan")

    def _model_forward(self, data):
        return self.model(data)                              # trainer.py:114

    def _graphed_step(self):
        """The CUDA-graph form of the step (lcrec_b200.train_graph), or None when this configuration keeps the eager loop:
        LCREC_TRAIN_GRAPH=0, a CPU device, an optimiser other than the fused Adam / AdamW, a subclass with its own forward
        plumbing (index_improve's EMA / reset steps read the host), active Dropout, cumulative-average BatchNorm."""
        if getattr(self, "_gstep", None) is not None:
            return self._gstep
        if getattr(self, "_gstep_checked", False):
            return None
        self._gstep_checked = True
        ok = (TRAIN_GRAPH and self.device.type == "cuda" and isinstance(self.optimizer, FusedAdam) and type(self) is Trainer
              and len(self.optimizer.param_groups) == 1
              and not any(isinstance(m, torch.nn.Dropout) and m.p > 0 for m in self.model.modules())
              and not any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.momentum is None for m in self.model.modules()))
        if ok:
            from .train_graph import GraphedTrainStep
            self._gstep = GraphedTrainStep(self.model, self.optimizer, forward_fn=self._model_forward, max_norm=1.0)
            return self._gstep
        return None

    def _train_epoch(self, train_data, epoch_idx):
        self.model.train()
        total_loss = 0
        total_recon_loss = 0
        bar = tqdm(train_data, total=len(train_data), ncols=100, desc=set_color(f"Train {epoch_idx}", "pink"))
        gstep = self._graphed_step()
        if gstep is not None:
            totals = [0.0, 0.0]

            def sink(loss, recon):
                totals[0] += loss
                totals[1] += recon
            for data in bar:
                gstep.step(data, self.device, sink)       # trainer.py:112-119 as one graph launch
                self.scheduler.step()
            gstep.flush(sink)
            return totals[0], totals[1]
        for data in bar:
            data = data.to(self.device)
            # the two status words the reference reads on the host inside the forward pass (vq.py:59, :81-82) are read at
            # the end of the step, where `loss.item()` synchronises anyway: the step is enqueued without draining the GPU
            with ops.defer_checks():
                self.optimizer.zero_grad()
                out, rq_loss, _ = self._model_forward(data)
                loss, loss_recon = self.model.compute_loss(out, rq_loss, xs=data)
                self._check_nan(loss)
                loss.backward()
                if isinstance(self.optimizer, FusedAdam) or hasattr(self.optimizer, "clip_and_step"):
                    self.optimizer.clip_and_step(1.0)                      # trainer.py:117-118 in one native call
                else:
                    torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1.0)
                    self.optimizer.step()
                self.scheduler.step()
                total_loss += loss.item()
                total_recon_loss += loss_recon.item()
        return total_loss, total_recon_loss

    @torch.no_grad()
    def _valid_epoch(self, valid_data):
        self.model.eval()
        bar = tqdm(valid_data, total=len(valid_data), ncols=100, desc=set_color("Evaluate   ", "pink"))
        chunks = []
        for data in bar:
            chunks.append(self.model.get_indices(data.to(self.device)).view(-1, len(self.model.rq.vq_layers)))
        codes = torch.cat(chunks, dim=0)
        num_sample = codes.shape[0]
        n_unique = ops.collisions(codes, [vq.n_e for vq in self.model.rq.vq_layers])["n_unique"]
        return (num_sample - n_unique) / num_sample

    def _save_checkpoint(self, epoch, collision_rate=1, ckpt_file=None):
        name = ckpt_file if ckpt_file else "epoch_%d_collision_%.4f_model.pth" % (epoch, collision_rate)
        ckpt_path = os.path.join(self.ckpt_dir, name)
        state = {"args": self.args, "epoch": epoch, "best_loss": self.best_loss,
                 "best_collision_rate": self.best_collision_rate, "state_dict": self.model.state_dict(),
                 "optimizer": self.optimizer.state_dict()}
        torch.save(state, ckpt_path, pickle_protocol=4)
        self.logger.info(set_color("Saving current", "blue") + f": {ckpt_path}")
        return ckpt_path

    def _generate_train_loss_output(self, epoch_idx, s_time, e_time, loss, recon_loss):
        head = (set_color("epoch %d training", "green") + " [" + set_color("time", "blue") + ": %.2fs, ") % (
            epoch_idx, e_time - s_time)
        return (head + set_color("train loss", "blue") + ": %.4f" % loss + ", " +
                set_color("reconstruction loss", "blue") + ": %.4f" % recon_loss + "]")

    def _generate_valid_output(self, epoch_idx, seconds, collision_rate):
        return (set_color("epoch %d evaluating", "green") + " [" + set_color("time", "blue") + ": %.2fs, " +
                set_color("collision_rate", "blue") + ": %f]") % (epoch_idx, seconds, collision_rate)

    def fit(self, data):
        for epoch_idx in range(self.epochs):
            t0 = time()
            train_loss, train_recon_loss = self._train_epoch(data, epoch_idx)
            self.logger.info(self._generate_train_loss_output(epoch_idx, t0, time(), train_loss, train_recon_loss))
            if (epoch_idx + 1) % self.eval_step != 0:
                continue
            t0 = time()
            collision_rate = self._valid_epoch(data)
            if train_loss < self.best_loss:
                self.best_loss = train_loss
                self._save_checkpoint(epoch=epoch_idx, ckpt_file=self.best_loss_ckpt)
            if collision_rate < self.best_collision_rate:
                self.best_collision_rate = collision_rate
                self._save_checkpoint(epoch_idx, collision_rate=collision_rate, ckpt_file=self.best_collision_ckpt)
            self.logger.info(self._generate_valid_output(epoch_idx, time() - t0, collision_rate))
            ckpt_path = self._save_checkpoint(epoch_idx, collision_rate=collision_rate)
            now_save = (-collision_rate, ckpt_path)
            if len(self.newest_save_queue) < self.save_limit:        # rotation, trainer.py:231-247
                self.newest_save_queue.append(now_save)
                heapq.heappush(self.best_save_heap, now_save)
            else:
                old_save = self.newest_save_queue.pop(0)
                self.newest_save_queue.append(now_save)
                if collision_rate < -self.best_save_heap[0][0]:
                    bad_save = heapq.heappop(self.best_save_heap)
                    heapq.heappush(self.best_save_heap, now_save)
                    if bad_save not in self.newest_save_queue:
                        delete_file(bad_save[1])
                if old_save not in self.best_save_heap:
                    delete_file(old_save[1])
        return self.best_loss, self.best_collision_rate
