"""Stand-alone CUDA-event timing of K4 (upsample + argmax + confusion matrix) at the eval shape.
    python profiles/time_k4.py [frames_per_launch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
C, h, w, H, W = 19, 128, 256, 1024, 2048
sets = []
for i in range(12):        # 12 x (labels 16.8 MB + logits 2.5 MB) per frame > L2 once n >= 1
    lg = torch.randn(n, C, h, w, device=dev, generator=g)
    lab = torch.randint(0, C, (n, H, W), device=dev, generator=g)
    lab[torch.rand(n, H, W, device=dev, generator=g) < 0.1] = 255
    sets.append((lg, lab))
cm = torch.zeros(C, C, dtype=torch.int64, device=dev)
for lg, lab in sets[:3]:
    _lib.upsample_argmax_confusion(lg, lab, (H, W), cm=cm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    for lg, lab in sets:
        _lib.upsample_argmax_confusion(lg, lab, (H, W), cm=cm)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * len(sets) * n)
alg = 8 * H * W + 4 * C * h * w + 8 * C * C
print(f"K4 dbg={os.environ.get('B200SEG_K4_DBG', '0')} frames/launch={n}: {us:.2f} us/frame  {alg / us / 1e3:.1f} GB/s algorithmic  cm_total={int(cm.sum())}")
