"""CUDA-event timing of the materialising align-corners upsample forward / backward (API-compat path), with
PREALLOCATED outputs (an allocation inside the timed loop costs more than the kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib

dev = torch.device("cuda", 0)
N, C, h, w, H, W = 8, 19, 64, 128, 512, 1024
lib = _lib.load()
x = torch.randn(N, C, h, w, device=dev)
outs = [torch.empty(N, C, H, W, device=dev) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream


def fwd(i):
    rc = lib.b200seg_upsample_bilinear_forward(x.data_ptr(), outs[i & 1].data_ptr(), N * C, h, w, H, W, 0, st)
    assert rc == 0


for i in range(3):
    fwd(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    fwd(i)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20
nbytes = N * C * H * W * 4
print(f"upsample fwd: {t * 1e3:.1f} us  {nbytes / t / 1e6:.0f} GB/s (write {nbytes / 1e6:.0f} MB)", flush=True)
go = [torch.randn(N, C, H, W, device=dev) for _ in range(2)]
gin = torch.empty(N, C, h, w, device=dev)


def bwd(i):
    rc = lib.b200seg_upsample_bilinear_backward(go[i & 1].data_ptr(), gin.data_ptr(), N * C, h, w, H, W, st)
    assert rc == 0


for i in range(3):
    bwd(i)
torch.cuda.synchronize()
e0.record()
for i in range(20):
    bwd(i)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20
print(f"upsample bwd: {t * 1e3:.1f} us  {nbytes / t / 1e6:.0f} GB/s (read {nbytes / 1e6:.0f} MB)", flush=True)
