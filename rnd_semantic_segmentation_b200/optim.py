"""Fused optimizer steps for the head / discriminator parameters.

``FusedSGD`` / ``FusedAdam`` ARE ``torch.optim.SGD`` / ``torch.optim.Adam`` (subclasses: same constructor, ``param_groups``,
``zero_grad``, ``state_dict`` / ``load_state_dict`` -- the ``optimizer_cls`` / ``optimizer_D`` entries of the reference's
checkpoints load unchanged, aspp_trainer.py:46-55, aspp_fada.py:29-40); only ``step()`` is replaced: one K8 launch per
parameter group instead of torch's chain of elementwise kernels, with the data-parallel mean (``grad_scale = 1/world_size``
after a SUM all-reduce) folded into the same pass.  The reference builds them as

    torch.optim.SGD(classifier.parameters(), lr=BASE_LR*10, momentum=MOMENTUM, weight_decay=WEIGHT_DECAY)   aspp_trainer.py:26
    torch.optim.Adam(model_D.parameters(), lr=BASE_LR_D, betas=(0.9, 0.99))                                   fada_adapter.py:24

and rewrites ``param_groups[i]['lr']`` every iteration from ``adjust_learning_rate`` (aspp_trainer.py:77-81), which keeps working
because the learning rate is read from the group at every step.  CUDA fp32 parameters only; no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["FusedSGD", "FusedAdam", "adjust_learning_rate"]


def adjust_learning_rate(method, base_lr, iters, max_iter, power):
    """core/utils/adapt_lr.py:12-17: poly schedule, NotImplementedError for anything else."""
    if method == 'poly':
        lr = base_lr * ((1 - float(iters) / max_iter) ** (power))
    else:
        raise NotImplementedError
    return lr


def _check_group_flags(group, who):
    for flag in ("maximize", "differentiable", "capturable", "amsgrad"):
        if group.get(flag):
            raise _lib.B200SegError(f"{who}: {flag}=True is not supported by the fused step")


_F32 = torch.float32


def _collect(group, who):
    """Parameters of the group that have a gradient, and those gradients as contiguous fp32 tensors (cheap per-step checks: the
    step is launch-bound, so the host side is kept to a few attribute reads per tensor)."""
    params, grads = [], []
    for p in group["params"]:
        g = p.grad
        if g is None:
            continue
        if not p.is_cuda:
            raise _lib.B200SegError(f"{who}: expected CUDA parameters (b200seg has no CPU fallback), got {p.device}")
        if p.dtype is not _F32 or not p.is_contiguous():
            raise _lib.B200SegError(f"{who}: parameters must be contiguous fp32 tensors")
        if g.dtype is not _F32 or not g.is_contiguous() or g.device != p.device:
            if g.is_sparse:
                raise _lib.B200SegError(f"{who}: sparse gradients are not supported")
            g = g.to(device=p.device, dtype=_F32).contiguous()
        params.append(p)
        grads.append(g)
    return params, grads


def _ptrs(tensors):
    return (_lib.c_vp * len(tensors))(*[t.data_ptr() for t in tensors])


def _launch(fn, device, *args):
    with _lib._on_device(device):
        _lib._check(fn(*args, _lib._stream()))


def _touched(params, *states):
    """The kernels write through raw pointers: tell autograd's version counters, so that whatever keys a cache on
    ``tensor._version`` (the modules' packed bf16 weights do) or saved a tensor for backward sees the in-place update."""
    torch.autograd.graph.increment_version(params)
    for st in states:
        if st:
            torch.autograd.graph.increment_version(st)


class _PlanMixin:
    """Per-group launch plans: the state tensors of a group are stable objects between ``load_state_dict`` calls, so their
    pointer tables (and the element counts) are built once; a step then costs one pass over the gradients on the host."""

    def _plans_init(self):
        self._plans = {}

    def load_state_dict(self, state_dict):
        self._plans = {}
        return super().load_state_dict(state_dict)

    def __getstate__(self):
        state = super().__getstate__()
        state["grad_scale"] = getattr(self, "grad_scale", 1.0)
        return state

    def __setstate__(self, state):                      # unpickled / deep-copied optimizers carry no plans (and no raw pointers)
        super().__setstate__(state)
        self._plans = {}
        if not hasattr(self, "grad_scale"):
            self.grad_scale = 1.0

    def _plan(self, gi, params, has_state=True):
        plan = self._plans.get(gi)
        if (plan is not None and plan.get("has_state", True) == has_state and len(plan["ids"]) == len(params)
                and all(a == id(b) for a, b in zip(plan["ids"], params))):
            return plan
        return None


def _state_ok(b, p):
    return b.dtype is _F32 and b.is_contiguous() and b.device == p.device and b.numel() == p.numel()


class FusedSGD(_PlanMixin, torch.optim.SGD):
    """torch.optim.SGD whose ``step()`` is one ``b200seg_sgd_step`` launch per parameter group.  ``grad_scale`` multiplies
    every gradient as it is read (set it to 1/world_size when the gradients were SUM-all-reduced)."""

    def __init__(self, params, *args, grad_scale: float = 1.0, **kwargs):
        kwargs.pop("foreach", None)
        kwargs.pop("fused", None)
        super().__init__(params, *args, **kwargs)
        self.grad_scale = float(grad_scale)
        self._plans_init()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            _check_group_flags(group, "FusedSGD")
            params, grads = _collect(group, "FusedSGD")
            if not params:
                continue
            momentum = float(group["momentum"])
            hyper = (float(group["lr"]), momentum, float(group["dampening"]), float(group["weight_decay"]), 1 if group["nesterov"] else 0)
            plan = self._plan(gi, params, has_state=momentum != 0)
            if plan is not None:
                _launch(lib.b200seg_sgd_step, params[0].device, len(params), _ptrs(params), _ptrs(grads), plan["bufs"], plan["numels"],
                        *hyper, 0, self.grad_scale)
                _touched(params, plan["keep"])
                continue
            state = self.state
            is_fresh = [momentum != 0 and state[p].get("momentum_buffer") is None for p in params]
            for first in (True, False):                     # one launch unless parameters joined the group at different times
                idx = [i for i, f in enumerate(is_fresh) if f == first]
                if not idx:
                    continue
                ps, gs = [params[i] for i in idx], [grads[i] for i in idx]
                bufs = None
                if momentum != 0:
                    if first:
                        for p in ps:
                            state[p]["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    bufs = [state[p]["momentum_buffer"] for p in ps]
                    if not all(_state_ok(b, p) for b, p in zip(bufs, ps)):
                        raise _lib.B200SegError("FusedSGD: momentum buffers must be contiguous fp32 tensors on the parameter's device")
                numels = (_lib.c_i64 * len(ps))(*[p.numel() for p in ps])
                _launch(lib.b200seg_sgd_step, ps[0].device, len(ps), _ptrs(ps), _ptrs(gs), _ptrs(bufs) if bufs else None, numels,
                        *hyper, 1 if first else 0, self.grad_scale)
                _touched(ps, bufs)
            bufs = [state[p]["momentum_buffer"] for p in params] if momentum != 0 else None
            self._plans[gi] = dict(ids=[id(p) for p in params], has_state=momentum != 0, keep=bufs, bufs=_ptrs(bufs) if bufs else None,
                                   numels=(_lib.c_i64 * len(params))(*[p.numel() for p in params]))
        return loss


class FusedAdam(_PlanMixin, torch.optim.Adam):
    """torch.optim.Adam whose ``step()`` is one ``b200seg_adam_step`` launch per parameter group (state keys ``step``,
    ``exp_avg``, ``exp_avg_sq`` as torch keeps them, so state dicts interchange)."""

    def __init__(self, params, *args, grad_scale: float = 1.0, **kwargs):
        kwargs.pop("foreach", None)
        kwargs.pop("fused", None)
        super().__init__(params, *args, **kwargs)
        self.grad_scale = float(grad_scale)
        self._plans_init()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            _check_group_flags(group, "FusedAdam")
            if isinstance(group["lr"], torch.Tensor):
                raise _lib.B200SegError("FusedAdam: tensor learning rates are not supported")
            params, grads = _collect(group, "FusedAdam")
            if not params:
                continue
            beta1, beta2 = group["betas"]
            hyper = (float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]))
            plan = self._plan(gi, params)
            if plan is not None:
                plan["steps"].add_(1)                            # the per-parameter CPU `step` scalars are views of this vector
                plan["step"] += 1
                _launch(lib.b200seg_adam_step, params[0].device, len(params), _ptrs(params), _ptrs(grads), plan["m"], plan["v"],
                        plan["numels"], *hyper, plan["step"], self.grad_scale)
                _touched(params, *plan["keep"])
                continue
            by_step = {}
            for p, g in zip(params, grads):
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=_F32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append((p, g, st))
            for step, items in by_step.items():             # one launch unless parameters joined the group at different times
                ps = [it[0] for it in items]
                ms, vs = [it[2]["exp_avg"] for it in items], [it[2]["exp_avg_sq"] for it in items]
                if not all(_state_ok(m, p) and _state_ok(v, p) for m, v, p in zip(ms, vs, ps)):
                    raise _lib.B200SegError("FusedAdam: state tensors must be contiguous fp32 tensors on the parameter's device")
                numels = (_lib.c_i64 * len(ps))(*[p.numel() for p in ps])
                _launch(lib.b200seg_adam_step, ps[0].device, len(ps), _ptrs(ps), _ptrs([it[1] for it in items]), _ptrs(ms), _ptrs(vs),
                        numels, *hyper, step, self.grad_scale)
                _touched(ps, ms, vs)
            if len(by_step) == 1 and all(not self.state[p]["step"].is_cuda for p in params):
                ms, vs = [self.state[p]["exp_avg"] for p in params], [self.state[p]["exp_avg_sq"] for p in params]
                step = next(iter(by_step))
                steps = torch.full((len(params),), float(step), dtype=_F32)      # one add_ per step instead of one per parameter
                for i, p in enumerate(params):
                    self.state[p]["step"] = steps[i]
                self._plans[gi] = dict(ids=[id(p) for p in params], keep=(ms, vs), m=_ptrs(ms), v=_ptrs(vs), steps=steps, step=step,
                                       numels=(_lib.c_i64 * len(params))(*[p.numel() for p in params]))
        return loss
