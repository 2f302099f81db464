"""Stand-alone timing of the PixelDiscriminator conv stack (K6) at the adversarial config (N=4, 2048 -> 256 -> 128 -> 2x19,
64x128): per-layer CUDA-event times through the library's profiling hooks + whole fwd / fwd+bwd, with the cuDNN fp32 (TF32 off
and on) stack beside it.  python profiles/time_disc.py [N]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib
from oracle import torch_oracle as to

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
C, h, w = 19, 64, 128
torch.manual_seed(0)
D = b200.PixelDiscriminator(2048, 256, num_classes=C).cuda()
ref = to.PixelDiscriminatorOracle(2048, 256, num_classes=C).cuda()
x = torch.relu(torch.randn(N, 2048, h, w, device="cuda"))
xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
go = torch.randn(N, 2 * C, h, w, device="cuda") * 1e-3


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def fwd(mod, inp):
    with torch.no_grad():
        return mod(inp)


def fwdbwd(mod, inp, need_x):
    inp = inp.detach().requires_grad_(need_x)
    for p in mod.parameters():
        p.grad = None
    mod(inp).backward(go)


_lib.conv_set_pair(os.environ.get("B200SEG_CONV_PAIR", "1") != "0")
print("conv pair (cta_group::2):", os.environ.get("B200SEG_CONV_PAIR", "1") != "0")
import rnd_semantic_segmentation_b200 as _b
_b.set_feature_pack_cache(0)
P = N * h * w
flops_fwd = 2 * 9 * P * (2048 * 256 + 256 * 128 + 128 * 2 * C)
print(f"N={N} P={P} fwd GFLOP={flops_fwd / 1e9:.1f}")
for name, inp in (("fp32 NCHW in", x), ("bf16 NHWC in (seam)", xb)):
    t = timeit(lambda: fwd(D, inp))
    print(f"ours {name}: fwd {t:.3f} ms = {flops_fwd / t / 1e9:.0f} TFLOP/s")
    t = timeit(lambda: fwdbwd(D, inp, True))
    print(f"ours {name}: fwd+bwd(dX,dW) {t:.3f} ms = {3 * flops_fwd / t / 1e9:.0f} TFLOP/s")
    t = timeit(lambda: fwdbwd(D, inp, False))
    print(f"ours {name}: fwd+bwd(dW only) {t:.3f} ms")
_lib.profile_enable(True)
for _ in range(5):
    fwdbwd(D, x, True)
torch.cuda.synchronize()
for k, (ms, n) in sorted(_lib.profile_read().items()):
    print(f"  {k}: {ms / n:.4f} ms x {n // 5} per step")
_lib.profile_enable(False)
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    t = timeit(lambda: fwd(ref, x), 5, 2)
    t2 = timeit(lambda: fwdbwd(ref, x, True), 5, 2)
    print(f"cuDNN fp32 (tf32={tf32}): fwd {t:.3f} ms, fwd+bwd {t2:.3f} ms")
