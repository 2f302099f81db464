"""Eval head forward (1 x 2048 x 128 x 256) and train head forward (8 x 2048 x 64 x 128) with 256-column tiles only vs the
per-problem tile width (B200SEG narrow tiles): per-kernel CUDA-event times.   python profiles/time_eval_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib, synth
RATES = [6, 12, 18, 24]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(2048, RATES, RATES, 19).to(dev).eval()
b200.set_feature_pack_cache(0)
for name, shape in (("eval 1x2048x128x256", (1, 2048, 128, 256)), ("train 8x2048x64x128", (8, 2048, 64, 128))):
    x = synth.make_features(*shape, device=dev)
    for narrow, sharing in ((0, 5), (1, 5), (1, 1), (1, 5), (1, 1)):
        _lib.gemm_set_narrow_tiles(narrow)
        _lib.gemm_set_sharing(sharing)
        with torch.no_grad():
            for _ in range(3):
                ref = head.logits(x)
            torch.cuda.synchronize()
            _lib.profile_enable(True)
            for _ in range(10):
                out = head.logits(x)
            torch.cuda.synchronize()
        prof = _lib.profile_read(); _lib.profile_enable(False)
        print(name, "narrow", narrow, "sharing", sharing, {k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items()}, "max|out|", float(out.abs().max()))
    del x
