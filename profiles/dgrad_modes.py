"""A/B of the head's fp32 NCHW data-gradient GEMM (b200seg_gemm_set_dgrad_mode): CUDA-event time of the kernel per mode, the
whole step per mode, and the difference of dX against mode 0.   python profiles/dgrad_modes.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth, _lib

RATES = [6, 12, 18, 24]
name = sys.argv[1] if len(sys.argv) > 1 else "train_b8_512x1024"
n, cin, h, w, H, W, C = synth.WORKLOADS[name]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev)
labels = synth.make_labels(n, H, W, C, device=dev)
params = list(head.parameters())


def step():
    xg = x.detach().requires_grad_(True)
    for p in params:
        p.grad = None
    loss, _ = head.forward_loss(xg, labels)
    loss.backward()
    return xg.grad


ref = None
for rep in range(2):
    for mode in [int(m) for m in os.environ.get("MODES", "0,1,2,3").split(",")]:
        _lib.gemm_set_dgrad_mode(mode)
        for _ in range(5):
            g = step()
        torch.cuda.synchronize()
        if ref is None:
            ref = g.clone()
        diff = (g - ref).abs().max().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        _lib.profile_enable(True)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        d = prof.get("head_dgrad_gemm", (0, 1))
        print(f"rep {rep} mode {mode}: step {ms:.4f} ms, dgrad kernel {1e3 * d[0] / max(d[1], 1):.1f} us, max |dX - dX(mode 0)| = {diff:.3e} (max |dX| {ref.abs().max().item():.3e})", flush=True)
_lib.gemm_set_dgrad_mode(1)
