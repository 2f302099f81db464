"""Eval-frame head forward: pack + GEMM against the GEMM with in-kernel fp32 NCHW -> bf16 conversion (b200seg_gemm_set_fwd_convert).
   python profiles/time_fwd_convert.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth, _lib

RATES = [6, 12, 18, 24]
dev = torch.device("cuda", 0)
for name in ("eval_1024x2048", "train_b8_512x1024"):
    n, cin, h, w, H, W, C = synth.WORKLOADS[name]
    torch.manual_seed(0)
    head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev).eval()
    xs = [synth.make_features(n, cin, h, w, seed=i, device=dev) for i in range(3)]
    ref = None
    for rep in range(2):
        for on in (False, True):
            _lib.gemm_set_fwd_convert(on)
            with torch.no_grad():
                for i in range(6):
                    lg = head.logits(xs[i % 3])
                torch.cuda.synchronize()
                if ref is None:
                    ref = head.logits(xs[0]).clone()
                same = torch.equal(head.logits(xs[0]), ref)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(30):
                    head.logits(xs[i % 3])
                e1.record()
                torch.cuda.synchronize()
            print(f"{name} rep {rep} fwd_convert={on}: {e0.elapsed_time(e1) / 30 * 1e3:.1f} us per head forward, logits identical: {same}", flush=True)
_lib.gemm_set_fwd_convert(False)
