"""Stand-alone CUDA-event timing of K2 (upsample + CE forward + low-res gradient) at the bench shape.
    python profiles/time_k2.py [N]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
C, h, w, H, W = 19, 64, 128, 512, 1024
sets = []
for i in range(6):
    lg = torch.randn(n, C, h, w, device=dev, generator=g)
    lab = torch.randint(0, C, (n, H, W), device=dev, generator=g)
    lab[torch.rand(n, H, W, device=dev, generator=g) < 0.1] = 255
    sets.append((lg, lab))
_lib.profile_enable(True)
for rep in range(4):
    for lg, lab in sets:
        x = lg.clone().requires_grad_(True)
        loss = ops.upsample_cross_entropy(x, lab)
        loss.backward()
torch.cuda.synchronize()
prof = _lib.profile_read()
ms, cnt = prof["upsample_ce_main"]
alg = 16 * n * H * W + 12 * n * C * h * w
print(f"K2 N={n}: {ms / cnt * 1e3:.2f} us/launch  {alg / (ms / cnt) / 1e6:.1f} GB/s algorithmic  loss={loss.item():.5f}")
