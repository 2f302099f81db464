"""Device time of the per-step weight packs (the weights change every training step): head 4 x [19,2048,3,3], discriminator
[256,2048,3,3] / [128,256,3,3] / 2 x [19,128,3,3].   python profiles/time_pack.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib, ops

dev = torch.device("cuda", 0)


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()                       # graph replay: device time without launch gaps
    with torch.cuda.graph(g):
        fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


hw = [torch.randn(19, 2048, 3, 3, device=dev) * 0.01 for _ in range(4)]
hb = [torch.randn(19, device=dev) for _ in range(4)]
print(f"head pack (Wp, WpT, bias_sum): {timeit(lambda: _lib.aspp_pack_weights(hw, hb)):.1f} us")
w1, w2 = torch.randn(256, 2048, 3, 3, device=dev) * 0.01, torch.randn(128, 256, 3, 3, device=dev) * 0.01
c1, c2 = torch.randn(19, 128, 3, 3, device=dev) * 0.01, torch.randn(19, 128, 3, 3, device=dev) * 0.01
b = torch.randn(19, device=dev)
print(f"conv pack layer 1 [256,2048]: {timeit(lambda: _lib.conv3x3_pack_weights([w1])):.1f} us")
print(f"conv pack layer 2 [128,256]:  {timeit(lambda: _lib.conv3x3_pack_weights([w2])):.1f} us")
print(f"conv pack cls1|cls2 [38,128]: {timeit(lambda: _lib.conv3x3_pack_weights([c1, c2])):.1f} us")
print(f"whole discriminator pack:     {timeit(lambda: ops.pack_discriminator_weights(w1, w2, c1, c2, b, b)):.1f} us")
