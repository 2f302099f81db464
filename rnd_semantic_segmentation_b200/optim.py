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


def _collect(group, who):
    params, grads = [], []
    for p in group["params"]:
        if p.grad is None:
            continue
        if not p.is_cuda:
            raise _lib.B200SegError(f"{who}: expected CUDA parameters (b200seg has no CPU fallback), got {p.device}")
        if p.grad.is_sparse:
            raise _lib.B200SegError(f"{who}: sparse gradients are not supported")
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise _lib.B200SegError(f"{who}: parameters must be contiguous fp32 tensors")
        g = p.grad
        if g.dtype != torch.float32 or not g.is_contiguous():
            g = g.float().contiguous()
        params.append(p)
        grads.append(g)
    return params, grads


class FusedSGD(torch.optim.SGD):
    """torch.optim.SGD whose ``step()`` is one ``b200seg_sgd_step`` launch per parameter group.  ``grad_scale`` multiplies
    every gradient as it is read (set it to 1/world_size when the gradients were SUM-all-reduced)."""

    def __init__(self, params, *args, grad_scale: float = 1.0, **kwargs):
        kwargs.pop("foreach", None)
        kwargs.pop("fused", None)
        super().__init__(params, *args, **kwargs)
        self.grad_scale = float(grad_scale)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            _check_group_flags(group, "FusedSGD")
            params, grads = _collect(group, "FusedSGD")
            if not params:
                continue
            momentum = float(group["momentum"])
            is_fresh = [momentum != 0 and self.state[p].get("momentum_buffer") is None for p in params]
            for first in (True, False):                     # one launch unless parameters joined the group at different times
                idx = [i for i, f in enumerate(is_fresh) if f == first]
                if not idx:
                    continue
                bufs = None
                if momentum != 0:
                    if first:
                        for i in idx:
                            self.state[params[i]]["momentum_buffer"] = torch.empty_like(params[i], memory_format=torch.contiguous_format)
                    bufs = [self.state[params[i]]["momentum_buffer"] for i in idx]
                _lib.sgd_step([params[i].data for i in idx], [grads[i] for i in idx], bufs, lr=float(group["lr"]), momentum=momentum,
                              dampening=float(group["dampening"]), weight_decay=float(group["weight_decay"]),
                              nesterov=bool(group["nesterov"]), first_step=first, grad_scale=self.grad_scale)
        return loss


class FusedAdam(torch.optim.Adam):
    """torch.optim.Adam whose ``step()`` is one ``b200seg_adam_step`` launch per parameter group (state keys ``step``,
    ``exp_avg``, ``exp_avg_sq`` as torch keeps them, so state dicts interchange)."""

    def __init__(self, params, *args, grad_scale: float = 1.0, **kwargs):
        kwargs.pop("foreach", None)
        kwargs.pop("fused", None)
        super().__init__(params, *args, **kwargs)
        self.grad_scale = float(grad_scale)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            _check_group_flags(group, "FusedAdam")
            if isinstance(group["lr"], torch.Tensor):
                raise _lib.B200SegError("FusedAdam: tensor learning rates are not supported")
            params, grads = _collect(group, "FusedAdam")
            if not params:
                continue
            by_step = {}
            for p, g in zip(params, grads):
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                by_step.setdefault(int(st["step"].item()), []).append((p, g, st))
            beta1, beta2 = group["betas"]
            for step, items in by_step.items():             # one launch unless parameters joined the group at different times
                _lib.adam_step([p.data for p, _, _ in items], [g for _, g, _ in items], [st["exp_avg"] for _, _, st in items],
                               [st["exp_avg_sq"] for _, _, st in items], step=step, lr=float(group["lr"]), beta1=float(beta1),
                               beta2=float(beta2), eps=float(group["eps"]), weight_decay=float(group["weight_decay"]),
                               grad_scale=self.grad_scale)
        return loss
