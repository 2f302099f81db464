"""Kernel-only CUDA-event timing of K3 (soft-label CE fwd / bwd) at the adversarial shape [4, 38, 512, 1024]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib, ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
N, K, H, W = 4, 38, 512, 1024
sets = [(torch.randn(N, K, H, W, device=dev, generator=g), torch.softmax(torch.randn(N, K, H, W, device=dev, generator=g), 1)) for _ in range(2)]
_lib.profile_enable(True)
for it in range(8):
    pred, soft = sets[it & 1]
    pr = pred.requires_grad_(True)
    pr.grad = None
    ops.soft_label_cross_entropy(pr, soft).backward()
torch.cuda.synchronize()
prof = _lib.profile_read()
px = N * H * W
for name, nbytes in (("soft_ce_fwd", 8 * K * px), ("soft_ce_bwd", 12 * K * px)):
    ms, cnt = prof[name]
    print(f"{name}: {ms / cnt * 1e3:.1f} us  {nbytes / (ms / cnt) / 1e6:.0f} GB/s")

# isolated: the C-ABI calls back to back, CUDA events around 10 calls
_lib.profile_enable(False)
for name, fn in (("fwd(stats)", lambda p, s: _lib.soft_ce_forward(p, s, None, want_stats=True)),
                 ("fwd(no stats)", lambda p, s: _lib.soft_ce_forward(p, s, None))):
    for it in range(3):
        fn(*sets[it & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10):
        fn(sets[it & 1][0].detach(), sets[it & 1][1])
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us/call")
loss, stats = _lib.soft_ce_forward(sets[0][0].detach(), sets[0][1], None, want_stats=True)
go = torch.ones(1, device=dev)
for it in range(3):
    _lib.soft_ce_backward(sets[0][0].detach(), sets[0][1], None, go, stats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    _lib.soft_ce_backward(sets[0][0].detach(), sets[0][1], None, go, stats)
e1.record()
torch.cuda.synchronize()
print(f"bwd(stats): {e0.elapsed_time(e1) / 10 * 1e3:.1f} us/call")
