"""ASPP_Classifier_V2 -- drop-in for core/models/classifiers/aspp/classifier.py:6-32.

Same constructor arguments, same ``forward(x, size=None)`` contract, same parameter owners
(``conv2d_list.{i}.weight/bias`` as real ``nn.Conv2d`` children, created in the same order with
the same init) so checkpoints, optimizer state and seeded initialisation are interchangeable with
the reference.  The arithmetic is NOT the four cuDNN convolutions: it is the fused tap-packed
tcgen05 GEMM head in csrc/aspp_head.cu (bf16 operands, fp32 accumulate).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import ops


class ASPP_Classifier_V2(nn.Module):
    def __init__(self, in_channels, dilation_series, padding_series, num_classes):
        super().__init__()
        self.conv2d_list = nn.ModuleList()
        for dilation, padding in zip(dilation_series, padding_series):
            self.conv2d_list.append(
                nn.Conv2d(in_channels, num_classes, kernel_size=3, stride=1, padding=padding, dilation=dilation, bias=True))
        for m in self.conv2d_list:                          # classifier.py:23-24
            m.weight.data.normal_(0, 0.01)
        self._packed = None
        self._packed_key = None
        # forward(x, size) returns a lazy.LazyLogits (fused head + upsample + loss when the caller's own criterion consumes it;
        # materialised on any other use).  False: always the materialised tensor (needed under torch's DDP wrapper).
        self.lazy = None        # None = automatic (lazy.lazy_enabled): lazy unless a multi-rank process group is initialised

    def out_channels_lowres(self):
        return int(self.conv2d_list[0].out_channels)

    # ---- helpers -------------------------------------------------------------------------
    def _rates(self):
        rates = []
        for m in self.conv2d_list:
            d, p = m.dilation, m.padding
            if d[0] != d[1] or tuple(p) != tuple(d) or m.kernel_size != (3, 3) or m.stride != (1, 1):
                raise NotImplementedError(
                    "b200seg ASPP head supports 3x3 / stride 1 / padding == dilation branches only "
                    f"(got dilation={d}, padding={p}); the reference's build_classifier always satisfies this")
            rates.append(int(d[0]))
        return rates

    def invalidate_packed(self):
        """Drop the cached bf16 weight pack (call after writing the parameters through ``.data``, a raw pointer or a
        CUDA-graph replay while the module is in eval mode)."""
        self._packed = None
        self._packed_key = None

    def _packed_weights(self):
        """Packed bf16 weights.  In training mode they are re-packed on EVERY call (one ~6 us kernel): parameter version
        counters miss in-place writes through ``.data`` -- the reference's own init idiom, classifier.py:23-24 -- and raw-pointer
        / CUDA-graph writers, and a stale pack would silently train on old weights.  Only eval mode reuses the pack, keyed
        on (data_ptr, _version); ``invalidate_packed()`` drops it explicitly."""
        params = [m.weight for m in self.conv2d_list] + [m.bias for m in self.conv2d_list]
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self.training or self._packed is None or key != self._packed_key:
            from . import _lib
            self._packed = _lib.aspp_pack_weights([m.weight.detach() for m in self.conv2d_list],
                                                  [m.bias.detach() for m in self.conv2d_list])
            self._packed_key = key
        return self._packed

    def logits(self, x: torch.Tensor) -> torch.Tensor:
        """Low-resolution logits [N,C,h,w] (classifier.py:27-29)."""
        return ops.aspp_head(x, [m.weight for m in self.conv2d_list], [m.bias for m in self.conv2d_list], self._rates(),
                             packed=self._packed_weights())

    # ---- reference API -------------------------------------------------------------------
    def forward(self, x, size=None):
        from . import lazy as _lazy
        if size is not None and x.is_cuda and _lazy.lazy_enabled(self.lazy):    # classifier.py:30-31, evaluated by whoever consumes it (lazy.py)
            from .lazy import LazyLogits, _LogitsSource
            return LazyLogits(_LogitsSource(self, x, size))
        out = self.logits(x)
        if size is not None:                                # classifier.py:30-31
            out = ops.upsample_bilinear_align_corners(out, size)
        return out

    # ---- fused extension (same result as criterion(self(x, size) / T, labels)) --------------
    def forward_loss(self, x, labels, ignore_index: int = 255, temperature: float = 1.0, grad_bucket=None):
        """loss = CrossEntropyLoss(ignore_index)(self(x, labels.shape[-2:]) / temperature, labels) with the
        upsample fused into the loss (aspp_trainer.py:88-91 / aspp_fada.py:91-95).  Returns (loss, detached low-res logits);
        gradients flow through the loss only.  ``grad_bucket``: a distributed.HeadGradBucket built over this module -- the
        parameter gradients then land in the bucket and its mean all-reduce overlaps the data-gradient GEMM."""
        return ops.aspp_head_loss(x, labels, [m.weight for m in self.conv2d_list], [m.bias for m in self.conv2d_list],
                                  self._rates(), ignore_index, temperature, grad_bucket=grad_bucket)
