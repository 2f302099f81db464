"""The short program profiled for the forward GEMM with in-kernel conversion: eval head forward at 1 x 2048 x 128 x 256.
    ncu --set full --clock-control none --import-source on -k regex:gemm_fwd_convert -s 2 -c 1 -o gpurun_out/prof_fwdx python profiles/prof_fwdx.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth

RATES = [6, 12, 18, 24]
dev = torch.device("cuda", 0)
head = synth.scale_head_for_unit_logits(b200.ASPP_Classifier_V2(2048, RATES, RATES, 19)).to(dev).eval()
xs = [synth.make_features(1, 2048, 128, 256, seed=5 + i, device=dev) for i in range(2)]
for i in range(4):
    with torch.no_grad():
        lg = head.logits(xs[i & 1])
torch.cuda.synchronize()
print(float(lg.sum()))
