"""Lazy stand-ins for the full-resolution tensors the reference's trainers pass between the modules and the losses
(SURVEY.md section 7 H4), so that UNMODIFIED trainer code reaches the fused kernels:

    output = classifier(feature_extractor(x), size)              # aspp_trainer.py:89      -> LazyLogits (nothing computed yet)
    loss = torch.nn.CrossEntropyLoss(ignore_index=255)(output, y) # aspp_trainer.py:61,91   -> ONE fused head + upsample + CE op
    loss.backward()                                               # aspp_trainer.py:92      -> fused backward (K2 + K1 dgrad / wgrad)

and, for the adversarial iteration (aspp_fada.py:91-125),

    src_pred = classifier(src_fea, size).div(1.8)                 # LazyLogits, temperature 1.8
    loss_seg = criterion(src_pred, src_label)                     # fused, caches the low-res logits
    soft = F.softmax(src_pred, dim=1).detach(); soft[soft > 0.9] = 0.9          # LazySoftLabel(clamp 0.9) on the cached logits
    D_pred = model_D(fea, size)                                   # LazyLogits of the discriminator
    soft_label_cross_entropy(D_pred, torch.cat((soft, torch.zeros_like(soft)), 1))   # -> fused K6 + K5 (slot 0; (zeros, soft): slot 1)

None of the [N,C,H,W] / [N,2C,H,W] tensors (319 MB each at the adversarial config) is ever written.  Every OTHER use of a lazy
object -- arithmetic, indexing, any torch.* function, any tensor method -- materialises the real tensor with the reference's
arithmetic (materialising upsample kernel, ATen softmax / clamp / cat) and carries on, autograd included, so semantics are those
of the reference whatever the caller does; only speed depends on the call pattern.

A lazy object is NOT a torch.Tensor instance.  Under torch's DistributedDataParallel wrapper (train_distill.py:54-62) the
module must return real tensors (DDP searches the outputs for autograd edges): set ``module.lazy = False`` there and use the
gradient buckets of ``distributed.py``.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops

__all__ = ["LazyLogits", "LazySoftLabel", "LazyZeros", "LazySlotLabel", "is_lazy", "materialize"]


def lazy_enabled(flag) -> bool:
    """Whether ``module(x, size)`` hands out a lazy object.  ``flag`` is the module's ``lazy`` attribute: True / False force it;
    None (the default) = automatic: env B200SEG_LAZY=0/1 if set, else lazy UNLESS a multi-rank process group is initialised.
    The reference wraps its modules in ``DistributedDataParallel(..., find_unused_parameters=True)`` (train_distill.py:54-62): DDP then
    walks the module OUTPUT for tensors to discover the used parameters, finds none in a lazy stand-in, and would mark every
    parameter unused.  Under a process group the default is therefore the materialised tensor (round-1 behaviour); the fused path
    is reached there through ``forward_loss`` / the gradient buckets (distributed.py), or by setting ``module.lazy = True`` when
    the module is not wrapped in DDP."""
    if flag is not None:
        return bool(flag)
    env = os.environ.get("B200SEG_LAZY")
    if env is not None:
        return env != "0"
    return not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)


def is_lazy(obj) -> bool:
    return isinstance(obj, _Lazy)


def materialize(obj):
    return obj.materialize() if isinstance(obj, _Lazy) else obj


class _Lazy:
    """Common tensor-like behaviour: shape queries are answered without computing; anything else materialises."""

    _shape: torch.Size
    _full = None

    # ---- cheap metadata ------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    def size(self, dim=None):
        return self._shape if dim is None else self._shape[dim]

    def dim(self):
        return len(self._shape)

    ndim = property(lambda self: len(self._shape))

    def numel(self):
        n = 1
        for s in self._shape:
            n *= int(s)
        return n

    @property
    def dtype(self):
        return torch.float32

    @property
    def is_cuda(self):
        return True

    def __len__(self):
        return int(self._shape[0])

    # ---- everything else: the real tensor ------------------------------------------------
    def materialize(self) -> torch.Tensor:
        raise NotImplementedError

    def __getattr__(self, name):              # tensor methods / attributes not defined on the lazy object
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        handler = _HANDLERS.get(func)
        if handler is not None:
            out = handler(*args, **kwargs)
            if out is not NotImplemented:
                return out
        args = _tree_materialize(args)
        kwargs = {k: _tree_materialize(v) for k, v in kwargs.items()}
        return func(*args, **kwargs)

    def __getitem__(self, idx):
        return self.materialize()[materialize(idx)]

    def __repr__(self):
        return f"{type(self).__name__}(shape={tuple(self._shape)})"


def _tree_materialize(v):
    if isinstance(v, _Lazy):
        return v.materialize()
    if isinstance(v, (list, tuple)):
        return type(v)(_tree_materialize(e) for e in v)
    return v


def _binary(name):
    def op(self, other):
        return getattr(self.materialize(), name)(materialize(other))
    op.__name__ = name
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__rtruediv__", "__pow__", "__lt__", "__le__",
           "__gt__", "__ge__", "__eq__", "__ne__", "__matmul__"):
    setattr(_Lazy, _n, _binary(_n))
_Lazy.__neg__ = lambda self: -self.materialize()
_Lazy.__hash__ = object.__hash__


# ------------------------------------------------------------------------------------------------
# logits:  module(x, size)  [/ temperature]
# ------------------------------------------------------------------------------------------------
class _LogitsSource:
    """Shared between a LazyLogits and its ``.div(t)`` descendants: the producing module and its input, and the low-res logits
    once some consumer has computed them."""
    __slots__ = ("module", "x", "size", "lowres", "lowres_tracked")

    def __init__(self, module, x, size):
        self.module, self.x, self.size = module, x, (int(size[0]), int(size[1]))
        self.lowres = None                    # [N,C,h,w]; detached when it came out of the fused loss op
        self.lowres_tracked = False           # True: ``lowres`` carries the autograd edge to x / the parameters

    def logits_lr(self, need_graph: bool) -> torch.Tensor:
        if self.lowres is None or (need_graph and not self.lowres_tracked):
            self.lowres = self.module.logits(self.x)
            self.lowres_tracked = torch.is_grad_enabled()
        return self.lowres


class LazyLogits(_Lazy):
    """``module(x, size)`` -- the align-corners upsampled logits [N,C,H,W] of an ASPP_Classifier_V2 or a PixelDiscriminator,
    optionally divided by a temperature -- before anyone looked at them."""

    def __init__(self, source: _LogitsSource, temperature: float = 1.0):
        self._src = source
        self._T = float(temperature)
        C = source.module.out_channels_lowres()
        self._shape = torch.Size((int(source.x.shape[0]), C) + source.size)

    @property
    def device(self):
        return self._src.x.device

    @property
    def requires_grad(self):
        return torch.is_grad_enabled() and (self._src.x.requires_grad or any(p.requires_grad for p in self._src.module.parameters()))

    # ---- lazy-preserving operations --------------------------------------------------------
    def div(self, other, *a, **k):
        if isinstance(other, (int, float)) and not a and not k and other != 0:
            return LazyLogits(self._src, self._T * float(other))          # aspp_fada.py:93-94 (src_pred.div(temperature))
        return self.materialize().div(materialize(other), *a, **k)

    __truediv__ = div
    true_divide = div

    def mul(self, other, *a, **k):
        if isinstance(other, (int, float)) and not a and not k and other != 0:
            return LazyLogits(self._src, self._T / float(other))
        return self.materialize().mul(materialize(other), *a, **k)

    def float(self):
        return self

    def contiguous(self, *a, **k):
        return self

    def detach(self):
        with torch.no_grad():
            lr = self._src.logits_lr(False).detach()
        src = _LogitsSource(_FixedLowres(lr), self._src.x.detach(), self._src.size)
        src.lowres, src.lowres_tracked = lr, True
        return LazyLogits(src, self._T)

    def softmax(self, dim=None, **k):
        return _softmax_handler(self, dim, **k)

    # ---- consumers -----------------------------------------------------------------------
    def cross_entropy(self, target, ignore_index: int = -100):
        """CrossEntropyLoss(ignore_index)(self, target) with the upsample fused into the loss (K2), and -- when the low-res
        logits have not been computed yet -- the head itself fused in front (``forward_loss``: no fp32 NCHW gradient tensor)."""
        src = self._src
        if src.lowres is None and hasattr(src.module, "forward_loss"):
            loss, lr = src.module.forward_loss(src.x, target, ignore_index=ignore_index, temperature=self._T)
            src.lowres, src.lowres_tracked = lr, False
            return loss
        return ops.upsample_cross_entropy(src.logits_lr(True), target, ignore_index, self._T)

    def materialize(self) -> torch.Tensor:
        if self._full is None:
            up = ops.upsample_bilinear_align_corners(self._src.logits_lr(torch.is_grad_enabled()), self._src.size)
            self._full = up if self._T == 1.0 else up.div(self._T)
        return self._full


class _FixedLowres(torch.nn.Module):
    """Stands in for the producing module of a detached LazyLogits: the low-res logits are a constant."""

    def __init__(self, lowres):
        super().__init__()
        self._lr = lowres

    def logits(self, x):
        return self._lr

    def out_channels_lowres(self):
        return int(self._lr.shape[1])


# ------------------------------------------------------------------------------------------------
# soft labels:  F.softmax(logits / T, dim=1).detach();  soft[soft > 0.9] = 0.9;  cat with zeros
# ------------------------------------------------------------------------------------------------
class LazySoftLabel(_Lazy):
    """``F.softmax(lazy_logits, dim=1)`` (aspp_fada.py:99,105) and what the reference does to it next: ``.detach()`` (:99,107),
    ``soft[soft > 0.9] = 0.9`` (:100,108)."""

    def __init__(self, logits: LazyLogits, detached: bool = False, clamp=None):
        self._logits, self._detached, self._clamp = logits, detached, clamp
        self._shape = logits.shape

    @property
    def device(self):
        return self._logits.device

    @property
    def requires_grad(self):
        return (not self._detached) and self._logits.requires_grad

    def detach(self):
        return LazySoftLabel(self._logits, True, self._clamp)

    def float(self):
        return self

    def __gt__(self, thr):
        if isinstance(thr, (int, float)):
            return _LazyGreater(self, float(thr))
        return self.materialize() > materialize(thr)

    def __setitem__(self, idx, value):
        # soft[soft > t] = t   ==   clamp(max=t): stays lazy.  (On a non-detached soft label the reference's in-place write
        # would also cut the gradient of the clamped entries; no reference caller does that, so that case materialises.)
        if (isinstance(idx, _LazyGreater) and idx.of is self and isinstance(value, (int, float)) and float(value) == idx.thr
                and self._detached and self._full is None):
            self._clamp = float(value) if self._clamp is None else min(self._clamp, float(value))
            return
        self.materialize()[materialize(idx)] = materialize(value)

    def materialize(self) -> torch.Tensor:
        if self._full is None:
            lg = self._logits
            if self._detached:
                with torch.no_grad():
                    soft = F.softmax(lg.detach().materialize(), dim=1)
            else:
                soft = F.softmax(lg.materialize(), dim=1)
            if self._clamp is not None:
                soft = torch.where(soft > self._clamp, torch.full_like(soft, self._clamp), soft)
            self._full = soft
        return self._full


class _LazyGreater(_Lazy):
    def __init__(self, of: LazySoftLabel, thr: float):
        self.of, self.thr = of, thr
        self._shape = of.shape

    def materialize(self):
        return self.of.materialize() > self.thr


class LazyZeros(_Lazy):
    """``torch.zeros_like(lazy_soft_label)`` (aspp_fada.py:111,120,124)."""

    def __init__(self, like: _Lazy):
        self._like = like
        self._shape = like.shape

    @property
    def device(self):
        return self._like.device

    def materialize(self):
        if self._full is None:
            self._full = torch.zeros(self._shape, dtype=torch.float32, device=self._like.device)
        return self._full


class LazySlotLabel(_Lazy):
    """``torch.cat((soft, zeros), 1)`` (source slot, aspp_fada.py:111,120) or ``torch.cat((zeros, soft), 1)`` (target slot, :124)."""

    def __init__(self, soft: LazySoftLabel, slot: int):
        self.soft, self.slot = soft, int(slot)
        n, c = soft.shape[:2]
        self._shape = torch.Size((n, 2 * c) + tuple(soft.shape[2:]))

    @property
    def device(self):
        return self.soft.device

    def float(self):
        return self

    def detach(self):
        return self

    def materialize(self):
        if self._full is None:
            s = self.soft.materialize()
            z = torch.zeros_like(s)
            self._full = torch.cat((s, z) if self.slot == 0 else (z, s), dim=1)
        return self._full


# ------------------------------------------------------------------------------------------------
# torch.* functions that keep things lazy (everything else materialises in _Lazy.__torch_function__)
# ------------------------------------------------------------------------------------------------
def _cross_entropy_handler(input, target, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean",
                           label_smoothing=0.0):
    if (isinstance(input, LazyLogits) and isinstance(target, torch.Tensor) and not target.is_floating_point() and weight is None
            and size_average is None and reduce is None and reduction == "mean" and label_smoothing == 0.0
            and target.dim() == 3 and tuple(target.shape[-2:]) == input._src.size):
        return input.cross_entropy(target, ignore_index)
    return NotImplemented


def _softmax_handler(input, dim=None, _stacklevel=3, dtype=None):
    if isinstance(input, LazyLogits) and dim in (1, -3) and dtype in (None, torch.float32):
        # the low-res logits are computed NOW (the reference evaluates the softmax at this line, e.g. before the optimizer step
        # of aspp_fada.py:114-115); only the full-resolution tensor stays virtual
        input._src.logits_lr(False)
        return LazySoftLabel(input)
    return NotImplemented


def _zeros_like_handler(input, **kwargs):
    if isinstance(input, (LazySoftLabel, LazyZeros)) and all(v is None for v in kwargs.values()):
        return LazyZeros(input)
    return NotImplemented


def _cat_handler(tensors, dim=0, **kwargs):
    if dim == 1 and not kwargs and isinstance(tensors, (list, tuple)) and len(tensors) == 2:
        a, b = tensors
        if isinstance(a, LazySoftLabel) and isinstance(b, LazyZeros) and b._like is a and a._detached:
            return LazySlotLabel(a, 0)
        if isinstance(a, LazyZeros) and isinstance(b, LazySoftLabel) and a._like is b and b._detached:
            return LazySlotLabel(b, 1)
    return NotImplemented


def _div_handler(input, other, **kwargs):
    if isinstance(input, LazyLogits) and not kwargs:
        return input.div(other)
    return NotImplemented


_HANDLERS = {
    F.cross_entropy: _cross_entropy_handler,
    F.softmax: _softmax_handler,
    torch.softmax: _softmax_handler,
    torch.zeros_like: _zeros_like_handler,
    torch.cat: _cat_handler,
    torch.div: _div_handler,
    torch.true_divide: _div_handler,
}


def soft_label_loss(pred, soft_label, pixel_weights=None):
    """The fused form of ``soft_label_cross_entropy(pred, soft_label)`` (utility.py:172-177) when ``pred`` is the lazy output of
    a PixelDiscriminator and ``soft_label`` a lazy slot label built from lazy head logits (aspp_fada.py:110-124): K6 conv stack
    + K5 loss tail on low-resolution tensors only.  Returns NotImplemented when the pattern does not apply."""
    if pixel_weights is not None or not isinstance(pred, LazyLogits) or not isinstance(soft_label, LazySlotLabel):
        return NotImplemented
    soft = soft_label.soft
    seg = soft._logits
    if pred._T != 1.0 or soft._full is not None or pred._src.size != seg._src.size or pred.shape[1] != 2 * seg.shape[1]:
        return NotImplemented
    with torch.no_grad():
        seg_lr = seg._src.logits_lr(False).detach()
    d_lr = pred._src.logits_lr(True)
    if tuple(d_lr.shape[-2:]) != tuple(seg_lr.shape[-2:]) or d_lr.shape[0] != seg_lr.shape[0]:
        return NotImplemented
    clamp = soft._clamp if soft._clamp is not None else float("inf")
    return ops.fada_soft_label_loss(d_lr, seg_lr, pred._src.size, soft_label.slot, temperature=seg._T, clamp=clamp)
