"""Short program for ncu: 2 warm-up + 1 profiled (cudaProfilerStart/Stop) train step and eval frame of the bench workloads.
    python profiles/prof_step.py [train_workload]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth

RATES = [6, 12, 18, 24]
wl = sys.argv[1] if len(sys.argv) > 1 else "train_b8_512x1024"
n, cin, h, w, H, W, C = synth.WORKLOADS[wl]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev)
labels = synth.make_labels(n, H, W, C, device=dev)
en, ecin, eh, ew, eH, eW, eC = synth.WORKLOADS["eval_1024x2048"]
ex = synth.make_features(1, ecin, eh, ew, seed=5, device=dev)
ey = synth.make_labels(1, eH, eW, eC, seed=6, device=dev)
ehead = synth.scale_head_for_unit_logits(b200.ASPP_Classifier_V2(ecin, RATES, RATES, eC)).to(dev).eval()
cm = torch.zeros(eC, eC, dtype=torch.int64, device=dev)
b200.set_feature_pack_cache(0)          # the same x every iteration: keep the fp32 -> bf16 pack in the capture
from rnd_semantic_segmentation_b200 import _lib
_lib.set_step_graphs(False)             # direct launches: every kernel is its own profiled launch
if os.environ.get("DGRAD_MODE"):
    _lib.gemm_set_dgrad_mode(int(os.environ["DGRAD_MODE"]))
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()            # ncu --profile-from-start off: only the third iteration is captured
    xg = x.detach().requires_grad_(True)
    for p in head.parameters():
        p.grad = None
    loss, _ = head.forward_loss(xg, labels)
    loss.backward()
    with torch.no_grad():
        lg = ehead.logits(ex)
    b200.segmentation_eval_step(lg, ey, cm=cm)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", loss.item(), int(cm.sum()))
