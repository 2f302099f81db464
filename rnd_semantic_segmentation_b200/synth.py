"""Synthetic Cityscapes-shaped inputs for tests and benchmarks.

There is no network for datasets or checkpoints: features are relu(N(0,1)) (layer4 output is post-ReLU),
head weights use the reference init (N(0, 0.01), default Conv2d bias), labels are uniform over the classes
with ``p_ignore`` of the pixels set to 255 in contiguous blocks (void regions).
"""
from __future__ import annotations

import torch

WORKLOADS = {
    # name: (batch, Cin, h, w, H, W, num_classes)  -- BASELINE.json configs
    "deeplabv2_r101_src": (2, 2048, 65, 129, 512, 1024, 19),
    "deeplabv2_r101_src_kvasir": (16, 2048, 44, 44, 352, 352, 2),
    "deeplabv2_r101_adv": (4, 2048, 64, 128, 512, 1024, 19),          # per domain
    "deeplabv2_r101_tgt_self_distill": (1, 2048, 64, 128, 512, 1024, 19),  # per GPU
    "train_b8_512x1024": (8, 2048, 64, 128, 512, 1024, 19),
    "eval_1024x2048": (1, 2048, 128, 256, 1024, 2048, 19),
}


def make_features(n, cin, h, w, seed=1234, device="cpu", dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.relu(torch.randn(n, cin, h, w, generator=g))
    return x.to(device=device, dtype=dtype)


def make_labels(n, H, W, num_classes, p_ignore=0.10, block=64, seed=1234, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed + 7919)
    lab = torch.randint(0, num_classes, (n, H, W), generator=g, dtype=torch.int64)
    if p_ignore > 0:
        bh, bw = (H + block - 1) // block, (W + block - 1) // block
        mask = torch.rand(n, bh, bw, generator=g) < p_ignore
        mask = mask.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W]
        lab[mask] = 255
    return lab.to(device)


def scale_head_for_unit_logits(head, factor=30.0):
    """Eval runs want logits with sigma ~ 1 (SURVEY.md section 8d cfg 5): scale the N(0, 0.01) init."""
    with torch.no_grad():
        for m in head.conv2d_list:
            m.weight.mul_(factor)
    return head
