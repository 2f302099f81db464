"""PixelDiscriminator -- drop-in for core/models/discriminator.py:31-50.

Parameter owners and state_dict keys (``D.0``, ``D.2``, ``cls1``, ``cls2``) are identical to the
reference.  The three 3x3 conv layers stay on the library path ("next"); what
is ours is the tail -- cat + align-corners upsample (+ the soft-label loss, see
``forward_soft_loss``).
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class PixelDiscriminator(nn.Module):
    def __init__(self, input_nc, ndf=512, num_classes=1):
        super().__init__()
        self.D = nn.Sequential(
            nn.Conv2d(input_nc, ndf, kernel_size=3, stride=1, padding=1),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Conv2d(ndf, ndf // 2, kernel_size=3, stride=1, padding=1),
            nn.LeakyReLU(negative_slope=0.2, inplace=True))
        self.cls1 = nn.Conv2d(ndf // 2, num_classes, kernel_size=3, stride=1, padding=1)
        self.cls2 = nn.Conv2d(ndf // 2, num_classes, kernel_size=3, stride=1, padding=1)

    def logits(self, x):
        mid = self.D(x)
        return torch.cat((self.cls1(mid), self.cls2(mid)), dim=1)       # discriminator.py:45-47

    def forward(self, x, size=None):
        out = self.logits(x)
        if size is not None:                                              # discriminator.py:48-49
            out = ops.upsample_bilinear_align_corners(out, size)
        return out

    # ---- fused extension ------------------------------------------------------------------
    def forward_soft_loss(self, x, seg_logits_lr, size, slot, temperature: float = 1.8, clamp: float = 0.9):
        """soft_label_cross_entropy(self(x, size), cat-slot(min(softmax(interpolate(seg_logits_lr, size) / T), clamp)))
        -- aspp_fada.py:110-111 (slot 0), :119-120 (slot 0), :123-124 (slot 1) -- without any full-resolution tensor."""
        return ops.fada_soft_label_loss(self.logits(x), seg_logits_lr, size, slot, temperature, clamp)
