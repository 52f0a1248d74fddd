"""Adam / AdamW with fused gradient clipping on the device: ``clip_grad_norm_(params, max_norm)`` + ``optimizer.step()`` of
reference ``index/trainer.py:117-119`` (optimisers built at :49-81) as two launches over all parameters
(``lcrec_adam_clip_step``).  A ``torch.optim.Optimizer`` subclass with the state layout of ``torch.optim.AdamW``
(``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so LR schedulers, ``state_dict()`` / ``load_state_dict()`` and the
checkpoints of the reference Trainer (:154-172) work unchanged in both directions.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _bump_versions(tensors) -> None:
    """The kernel writes the parameters through raw pointers; consumers that cache derived data by ``Tensor._version``
    (MLPLayers keeps split copies of its weights) must see an in-place update, exactly as after torch.optim's ``add_``."""
    try:
        torch._C._autograd._unsafe_set_version_counter(tuple(tensors), tuple(t._version + 1 for t in tensors))
    except (AttributeError, TypeError, RuntimeError):
        torch._foreach_add_(list(tensors), 0.0)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled_weight_decay=decoupled, amsgrad=False,
                        maximize=False, foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)

    # ---- device-resident per-step scalars (CUDA-graph replay of the training step)
    def prepare_hyper(self) -> None:
        """Advance every parameter's step count and publish this step's scalars {1 - lr wd, lr / (1 - beta1^t),
        sqrt(1 - beta2^t)} to the device (pinned ring -> 12-byte async copy on the current stream).  The following
        ``clip_and_step(..., device_hyper=True)`` - eager or replayed from a captured graph - reads them from there, so the
        learning-rate schedule and the bias corrections keep advancing although the launches are frozen in the graph."""
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdam.prepare_hyper: one parameter group expected")
        group = self.param_groups[0]
        ps = [p for p in group["params"] if p.requires_grad]
        dev = ps[0].device
        step = None
        for p in ps:
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            step = int(st["step"]) if step is None else step
        if getattr(self, "_hyper_dev", None) is None:
            self._hyper_dev = torch.zeros(3, dtype=torch.float32, device=dev)
            self._hyper_ring = torch.zeros((64, 3), dtype=torch.float32).pin_memory()
            self._hyper_slot = 0
        b1, b2 = group["betas"]
        out = (C.c_float * 3)()
        _lib.check(_lib.load().lcrec_adam_hyper(float(group["lr"]), float(b1), float(b2), float(group["weight_decay"]), step, out))
        slot = self._hyper_ring[self._hyper_slot % 64]
        self._hyper_slot += 1
        slot[0], slot[1], slot[2] = out[0], out[1], out[2]
        self._hyper_dev.copy_(slot, non_blocking=True)

    @torch.no_grad()
    def clip_and_step(self, max_norm: float = 0.0, want_norm: bool = False, device_hyper: bool = False):
        """Gradient clipping to ``max_norm`` (<= 0: none) over ALL parameter groups (like ``clip_grad_norm_`` on
        ``model.parameters()``) followed by the update.  Returns the total gradient norm (a device tensor) if asked."""
        self._opt_called = True          # what the step wrapper of an LR scheduler records (its call-order warning)
        groups = []
        all_grads = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam: contiguous fp32 CUDA parameters with dense gradients only "
                                       "(lcrec_b200 has no CPU fallback)")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            groups.append((group, ps))
            all_grads += [p.grad for p in ps]
        if not all_grads:
            return None
        lib = _lib.load()
        dev = all_grads[0].device
        norm_out = torch.zeros(1, dtype=torch.float32, device=dev) if (want_norm and max_norm > 0) else None
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            if max_norm > 0 and len(groups) > 1:
                # one norm over every group: clip first (pass 1 + scaling through a zero-lr call is not needed: groups
                # share the norm only if they are updated together, so merge groups with equal hyper-parameters)
                raise RuntimeError("FusedAdam.clip_and_step: one parameter group expected when clipping")
            for group, ps in groups:
                if not ps:
                    continue
                steps = set()
                if not device_hyper:
                    for p in ps:
                        st = self.state[p]
                        st["step"] += 1                              # a CPU scalar tensor, as torch.optim keeps it
                        steps.add(int(st["step"]))
                    if len(steps) != 1:
                        raise RuntimeError("FusedAdam: parameters of one group must share the step count")
                n = len(ps)
                numel = (C.c_int64 * n)(*[p.numel() for p in ps])
                ws_bytes = int(lib.lcrec_adam_workspace_bytes(n, numel))
                if getattr(self, "_ws", None) is None or self._ws.numel() < ws_bytes or self._ws.device != dev:
                    self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                b1, b2 = group["betas"]
                if device_hyper:                                  # scalars of this step were published by prepare_hyper()
                    _lib.check(lib.lcrec_adam_clip_step_dev(
                        n, _lib.ptr_array([p.data_ptr() for p in ps]), _lib.ptr_array([p.grad.data_ptr() for p in ps]),
                        _lib.ptr_array([self.state[p]["exp_avg"].data_ptr() for p in ps]),
                        _lib.ptr_array([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]), numel,
                        C.c_void_p(self._hyper_dev.data_ptr()), float(b1), float(b2), float(group["eps"]),
                        float(group["weight_decay"]), int(bool(group.get("decoupled_weight_decay", True))), float(max_norm), 1,
                        C.c_void_p(0 if norm_out is None else norm_out.data_ptr()), C.c_void_p(self._ws.data_ptr()),
                        self._ws.numel(), stream))
                    _bump_versions(ps)
                    continue
                _lib.check(lib.lcrec_adam_clip_step(
                    n, _lib.ptr_array([p.data_ptr() for p in ps]), _lib.ptr_array([p.grad.data_ptr() for p in ps]),
                    _lib.ptr_array([self.state[p]["exp_avg"].data_ptr() for p in ps]),
                    _lib.ptr_array([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]), numel, float(group["lr"]),
                    float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), int(bool(group.get("decoupled_weight_decay", True))),
                    steps.pop(), float(max_norm), 1, C.c_void_p(0 if norm_out is None else norm_out.data_ptr()),
                    C.c_void_p(self._ws.data_ptr()), self._ws.numel(), stream))
                _bump_versions(ps)
        return norm_out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.clip_and_step(0.0)
        return loss


class FusedSimple(torch.optim.Optimizer):
    """SGD / Adagrad / RMSprop as the reference's ``_build_optimizer`` constructs them (index/trainer.py:62-75: only ``lr`` and
    ``weight_decay`` are passed, everything else is the torch.optim default) with ``clip_grad_norm_`` fused in
    (``lcrec_simple_opt_clip_step``).  State keys follow torch.optim (``sum`` / ``square_avg`` + ``step``), so ``state_dict()``
    is interchangeable with ``torch.optim.Adagrad`` / ``RMSprop`` / ``SGD`` (no momentum buffer)."""
    KINDS = {"sgd": 1, "adagrad": 2, "rmsprop": 3}

    def __init__(self, params, kind: str, lr=1e-2, weight_decay=0.0):
        kind = kind.lower()
        if kind not in self.KINDS:
            raise ValueError(f"FusedSimple: unknown optimiser {kind!r}")
        defaults = dict(lr=lr, weight_decay=weight_decay)
        if kind == "adagrad":
            defaults.update(lr_decay=0, eps=1e-10, initial_accumulator_value=0)
        elif kind == "rmsprop":
            defaults.update(alpha=0.99, eps=1e-8, momentum=0, centered=False)
        else:
            defaults.update(momentum=0, dampening=0, nesterov=False)
        super().__init__(params, defaults)
        self.kind = kind
        self._state_key = {"sgd": None, "adagrad": "sum", "rmsprop": "square_avg"}[kind]

    @torch.no_grad()
    def clip_and_step(self, max_norm: float = 0.0):
        self._opt_called = True
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedSimple.clip_and_step: one parameter group expected")
        group = self.param_groups[0]
        ps = [p for p in group["params"] if p.grad is not None]
        if not ps:
            return
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                raise RuntimeError("FusedSimple: contiguous fp32 CUDA parameters with dense gradients only (no CPU fallback)")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            st = self.state[p]
            if self._state_key is not None and self._state_key not in st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st[self._state_key] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if "step" in st:
                st["step"] += 1
        lib = _lib.load()
        dev = ps[0].device
        n = len(ps)
        numel = (C.c_int64 * n)(*[p.numel() for p in ps])
        ws_bytes = int(lib.lcrec_adam_workspace_bytes(n, numel))
        if getattr(self, "_ws", None) is None or self._ws.numel() < ws_bytes or self._ws.device != dev:
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        state = None if self._state_key is None else _lib.ptr_array([self.state[p][self._state_key].data_ptr() for p in ps])
        with torch.cuda.device(dev):
            _lib.check(lib.lcrec_simple_opt_clip_step(
                self.KINDS[self.kind], n, _lib.ptr_array([p.data_ptr() for p in ps]), _lib.ptr_array([p.grad.data_ptr() for p in ps]), state,
                numel, float(group["lr"]), float(group["weight_decay"]), float(group.get("alpha", 0.99)), float(group.get("eps", 0.0)),
                float(max_norm), 1, C.c_void_p(0), C.c_void_p(self._ws.data_ptr()), self._ws.numel(),
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        _bump_versions(ps)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.clip_and_step(0.0)
        return loss
