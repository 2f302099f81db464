"""A/B of the head's forward GEMM (b200seg_gemm_set_fwd_mode): kernel time by CUDA events per mode and the difference of the
logits against mode 0.   python profiles/fwd_modes.py [workload]"""
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
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev).eval()
x = synth.make_features(n, cin, h, w, device=dev)
ref = None
for rep in range(2):
    for mode in (0, 1):
        _lib.gemm_set_fwd_mode(mode)
        with torch.no_grad():
            for _ in range(5):
                lg = head.logits(x)
            torch.cuda.synchronize()
            if ref is None:
                ref = lg.clone()
            diff = (lg - ref).abs().max().item()
            _lib.profile_enable(True)
            for _ in range(10):
                head.logits(x)
            torch.cuda.synchronize()
            prof = _lib.profile_read()
            _lib.profile_enable(False)
        d = prof.get("head_fwd_gemm", (0, 1))
        print(f"rep {rep} mode {mode}: forward GEMM {1e3 * d[0] / max(d[1], 1):.1f} us, max |logits - logits(mode 0)| = {diff:.3e} (max |logits| {ref.abs().max().item():.3e})", flush=True)
_lib.gemm_set_fwd_mode(1)
