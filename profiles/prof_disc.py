"""Short program for ncu: 2 warm-up + 1 profiled (cudaProfilerStart/Stop) PixelDiscriminator pass of the adversarial config
(N=4, 2048 -> 256 -> 128 -> 2 x 19 at 64 x 128): forward + fused soft-label loss + backward (dX, dW, db).
    python profiles/prof_disc.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth

n, cin, h, w, H, W, C = synth.WORKLOADS["deeplabv2_r101_adv"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
D = b200.PixelDiscriminator(cin, 256, num_classes=C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev)
seg = torch.randn(n, C, h, w, device=dev)
b200.set_feature_pack_cache(0)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    xg = x.detach().requires_grad_(True)
    for p in D.parameters():
        p.grad = None
    loss = D.forward_soft_loss(xg, seg, (H, W), slot=0)
    loss.backward()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", loss.item(), float(xg.grad.abs().max()))
