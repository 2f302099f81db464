"""CUDA-event timing of confusion_from_pred (confusion_matrix / intersectionAndUnionGPU on a materialised prediction map)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib

dev = torch.device("cuda", 0)
C, H, W = 19, 1024, 2048
g = torch.Generator(device=dev).manual_seed(0)
sets = []
for i in range(6):
    pd = torch.randint(0, C, (H, W), device=dev, generator=g)
    gt = torch.randint(0, C, (H, W), device=dev, generator=g)
    gt[torch.rand(H, W, device=dev, generator=g) < 0.1] = 255
    sets.append((pd, gt))
cm = torch.zeros(C, C, dtype=torch.int64, device=dev)
for pd, gt in sets[:3]:
    _lib.confusion_from_pred(pd, gt, C, cm=cm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, data in (("random labels", sets),):
    e0.record()
    for rep in range(5):
        for pd, gt in data:
            _lib.confusion_from_pred(pd, gt, C, cm=cm)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 30
    print(f"confusion_from_pred ({name}): {t * 1e3:.1f} us/frame  {16 * H * W / t / 1e6:.0f} GB/s")
# coherent (real-image-like) maps: 64x64 blocks
pdc = torch.randint(0, C, (H // 64, W // 64), device=dev, generator=g).repeat_interleave(64, 0).repeat_interleave(64, 1).contiguous()
e0.record()
for rep in range(30):
    _lib.confusion_from_pred(pdc, pdc, C, cm=cm)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 30
print(f"confusion_from_pred (blocky maps): {t * 1e3:.1f} us/frame  {16 * H * W / t / 1e6:.0f} GB/s")
# 8 frames per call: amortises the ~15 us Python/ctypes call overhead
pd8 = torch.cat([s[0].reshape(-1) for s in sets] + [sets[0][0].reshape(-1), sets[1][0].reshape(-1)])
gt8 = torch.cat([s[1].reshape(-1) for s in sets] + [sets[0][1].reshape(-1), sets[1][1].reshape(-1)])
for _ in range(3):
    _lib.confusion_from_pred(pd8, gt8, C, cm=cm)
torch.cuda.synchronize()
e0.record()
for rep in range(10):
    _lib.confusion_from_pred(pd8, gt8, C, cm=cm)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print(f"confusion_from_pred (8 frames per call): {t * 1e3 / 8:.1f} us/frame  {16 * pd8.numel() / t / 1e6:.0f} GB/s")
