"""K7: row-walking vs per-pixel kernel as a function of the number of ensemble members (1024x2048, members at 128x256)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib

C, H, W = 19, 1024, 2048
torch.manual_seed(0)
labels = torch.randint(0, C, (1, H, W), device="cuda")
cm = torch.zeros(C, C, dtype=torch.int64, device="cuda")


def timeit(fn, iters=20, warm=3):
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


for nm in (1, 2, 3, 4):
    members = [torch.randn(1, C, 128, 256, device="cuda") for _ in range(nm)]
    flips = [bool(k & 1) for k in range(nm)]
    res = {}
    for name, mode, fast in (("fast", True, True), ("rows", True, False), ("pixel", False, False)):
        _lib.tta_set_row_walk(mode, fast=fast)
        res[name] = timeit(lambda: _lib.tta_argmax_confusion(members, flips, (H, W), labels=labels, divisors=(nm,), cm=cm, want_pred=True))
    _lib.tta_set_row_walk(True)
    print(f"{nm} members: labels-only fast path {res['fast'] * 1e3:.1f} us, row-walking exact {res['rows'] * 1e3:.1f} us, "
          f"per-pixel {res['pixel'] * 1e3:.1f} us")
