"""The short program profiled for K2: a few fused upsample + CE forward/backward launches at the bench shape.
    ncu --set full --clock-control none --import-source on -k regex:upsample_ce_main -s 2 -c 1 -o gpurun_out/prof_k2 python profiles/prof_k2.py [variant] [u8]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib, ops

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
u8 = len(sys.argv) > 2 and sys.argv[2] == "u8"
_lib.upsample_ce_set_variant(variant)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n, C, h, w, H, W = 8, 19, 64, 128, 512, 1024
lg = 2.0 * torch.randn(n, C, h, w, device=dev, generator=g)
lab = torch.randint(0, C, (n, H, W), device=dev, generator=g)
blocks = torch.rand(n, H // 64, W // 64, device=dev, generator=g) < 0.1
lab[blocks.repeat_interleave(64, 1).repeat_interleave(64, 2)] = 255
if u8:
    lab = lab.to(torch.uint8)
for _ in range(4):
    x = lg.clone().requires_grad_(True)
    loss = ops.upsample_cross_entropy(x, lab)
    loss.backward()
torch.cuda.synchronize()
print(loss.item())
