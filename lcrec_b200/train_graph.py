"""The training step of reference ``index/trainer.py:111-123`` captured ONCE per batch size in a CUDA graph.

A step is ~60 kernel launches (encoder / decoder GEMMs forward and backward, operand splits, residual quantiser, dense
Sinkhorn, loss, clipping + AdamW); issued one by one from Python it is host-bound (4.35 ms at batch 1024 in round 1, most of
it launch gaps).  ``GraphedTrainStep`` runs the first steps of every batch size eagerly (k-means initialisation, allocator
warm-up), then records the SAME Python step - zero_grad, forward, loss, backward, clip + optimiser - with
``torch.cuda.graph`` and from then on replays it with one ``cudaGraphLaunch``:

* the batch is copied into a static input tensor (pinned host memory -> device, same stream) before the replay;
* the learning rate and Adam's bias corrections live in device memory (``FusedAdam.prepare_hyper`` + ``lcrec_adam_clip_step_dev``),
  so the LR schedule keeps advancing;
* the two losses and the status words the reference reads on the host inside its forward (``assert amplitude > 0``, vq.py:59;
  NaN print, vq.py:81-82) come back in ONE small pinned buffer per step; the host looks at step k's while step k + 1 is
  already running (``ValueError("Training loss is nan")`` of trainer.py:93-95 is raised one step late at most).

Nothing here computes: every launch inside the graph is a kernel of liblcrec_b200.so or a torch elementwise / copy.
Capture needs every launch of the step to be stream-ordered without host reads; if a configuration cannot be captured
(dropout, cumulative-average BatchNorm) the step stays eager and ``capture_error`` says why.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import ops
from .optim import FusedAdam, _bump_versions


class _Entry:
    def __init__(self):
        self.eager_left = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.x: Optional[torch.Tensor] = None
        self.report: Optional[torch.Tensor] = None
        self.kinds: List[str] = []
        self.n_launches = 0                  # kernels of liblcrec_b200.so recorded in the graph


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: FusedAdam, forward_fn: Optional[Callable] = None, max_norm: float = 1.0,
                 warmup: int = 3, enabled: bool = True, post_backward: Optional[Callable] = None):
        self.model, self.optimizer = model, optimizer
        self.forward_fn = forward_fn or (lambda x: model(x))
        self.max_norm, self.warmup = float(max_norm), int(warmup)
        self.post_backward = post_backward              # e.g. the gradient all-reduce of a data-parallel trainer
        self.entries: Dict[int, _Entry] = {}
        self.capture_error: Optional[str] = None if enabled else "disabled"
        self.replays = self.eager_steps = 0
        self.replayed_launches = 0           # library kernels executed through graph replays (the library's own counter only sees eager launches)
        self._ring = None
        self._slot = 0
        self._pending = []                              # (event, pinned row, kinds) of steps whose report is in flight
        self._versioned = [t for t in list(model.parameters()) + list(model.buffers()) if t.is_floating_point()]

    # ---- the step itself (the same code eagerly and under capture)
    def _body(self, x: torch.Tensor):
        self.optimizer.zero_grad(set_to_none=True)
        out, rq_loss, _ = self.forward_fn(x)
        loss, loss_recon = self.model.compute_loss(out, rq_loss, xs=x)
        loss.backward()
        if self.post_backward is not None:
            self.post_backward()
        self.optimizer.clip_and_step(self.max_norm, device_hyper=True)
        return loss.detach(), loss_recon.detach()

    def _run(self, x: torch.Tensor):
        """-> (report tensor [loss, recon, status words...], kinds of the status words)"""
        with ops.defer_checks():
            loss, recon = self._body(x)
            pend = ops.take_pending()
        words = [w.reshape(-1)[0].to(torch.float32) for _, w in pend]
        return torch.stack([loss.float(), recon.float(), *words]), [k for k, _ in pend]

    # ---- report plumbing: one pinned row per step, read one step late
    def _publish(self, report: torch.Tensor, kinds: List[str]) -> None:
        if self._ring is None or self._ring.shape[1] < report.numel():
            self._ring = torch.zeros((8, max(8, report.numel())), dtype=torch.float32).pin_memory()
        row = self._ring[self._slot % 8][: report.numel()]
        self._slot += 1
        row.copy_(report, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((ev, row, kinds))

    def _collect(self, keep: int, sink: Callable[[float, float], None]) -> None:
        while len(self._pending) > keep:
            ev, row, kinds = self._pending.pop(0)
            ev.synchronize()
            vals = row.tolist()
            for kind, v in zip(kinds, vals[2:]):
                ops.evaluate_check_value(kind, int(v))
            if vals[0] != vals[0]:
                raise ValueError("Training loss is nan")             # trainer.py:93-95
            sink(vals[0], vals[1])

    def flush(self, sink: Callable[[float, float], None]) -> None:
        self._collect(0, sink)

    # ---- one training step on `data` (n, in_dim): CPU (pinned or pageable) or CUDA tensor
    def step(self, data: torch.Tensor, device: torch.device, sink: Callable[[float, float], None]) -> None:
        n = int(data.shape[0])
        e = self.entries.get(n)
        if e is None:
            e = self.entries[n] = _Entry()
            e.eager_left = self.warmup
        self.optimizer.prepare_hyper()
        pending_init = any(not getattr(q, "initted", True) for q in getattr(getattr(self.model, "rq", None), "vq_layers", []))
        if e.graph is None and (e.eager_left > 0 or pending_init or self.capture_error is not None):
            x = data.to(device, non_blocking=True)
            report, kinds = self._run(x)
            e.eager_left -= 1
            self.eager_steps += 1
        else:
            if e.graph is None:
                e.x = torch.empty((n,) + tuple(data.shape[1:]), dtype=torch.float32, device=device)
                e.x.copy_(data, non_blocking=True)
                try:
                    torch.cuda.synchronize(device)
                    g = torch.cuda.CUDAGraph()
                    before = ops.launch_count()
                    with torch.cuda.graph(g):
                        e.report, e.kinds = self._run(e.x)
                    e.n_launches = ops.launch_count() - before
                    e.graph = g
                except Exception as exc:  # noqa: BLE001 - fall back to the eager step, loudly
                    self.capture_error = f"{type(exc).__name__}: {exc}"
                    ops.take_pending()
                    torch.cuda.synchronize(device)
                    import warnings
                    warnings.warn(f"lcrec_b200: the training step could not be captured in a CUDA graph ({self.capture_error}); "
                                  "running it eagerly")
                    report, kinds = self._run(data.to(device, non_blocking=True))
                    self._publish(report, kinds)
                    self._collect(1, sink)
                    return
            else:
                e.x.copy_(data, non_blocking=True)
            e.graph.replay()
            _bump_versions(self._versioned)        # the replay rewrote parameters / running statistics behind torch's back
            self.replays += 1
            self.replayed_launches += e.n_launches
            report, kinds = e.report, e.kinds
        self._publish(report, kinds)
        self._collect(1, sink)
