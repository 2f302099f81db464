"""Short program for ncu: 2 warm-up + 1 profiled (cudaProfilerStart/Stop) iteration of K7 (flip test-time augmentation of one
1024x2048 frame) and K8 (head SGD step, discriminator Adam step).   python profiles/prof_tta_optim.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib

dev = torch.device("cuda", 0)
torch.manual_seed(0)
C, H, W, h, w = 19, 1024, 2048, 128, 256
members = [torch.randn(1, C, h, w, device=dev) for _ in range(2)]
labels = torch.randint(0, C, (1, H, W), device=dev)
cm = torch.zeros(C, C, dtype=torch.int64, device=dev)


def group(shapes):
    ps = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.01) for s in shapes]
    for p in ps:
        p.grad = torch.randn_like(p) * 0.01
    return ps


sgd = b200.FusedSGD(group([(19, 2048, 3, 3), (19,)] * 4), lr=2.5e-3, momentum=0.9, weight_decay=5e-4)
adam = b200.FusedAdam(group([(256, 2048, 3, 3), (256,), (128, 256, 3, 3), (128,), (19, 128, 3, 3), (19,), (19, 128, 3, 3), (19,)]),
                      lr=1e-4, betas=(0.9, 0.99))
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    _lib.tta_argmax_confusion(members, [False, True], (H, W), labels=labels, divisors=(2,), cm=cm)          # row-walking kernel
    _lib.tta_argmax_confusion(members + members[:1], [False, True, False], (H, W), labels=labels, divisors=(3,), cm=cm)   # per-pixel kernel
    sgd.step()
    adam.step()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", int(cm.sum()))
