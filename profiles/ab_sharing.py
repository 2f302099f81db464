"""A/B of the head GEMMs' operand-sharing modes (b200seg_gemm_set_sharing): 5 = 2-CTA multicast pairs only (the earlier default),
1 = cta_group::2 pairs for forward / weight gradient (current default); per-kernel CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib, synth
RATES = [6, 12, 18, 24]
n, cin, h, w, H, W, C = synth.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "train_b8_512x1024"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev)
labels = synth.make_labels(n, H, W, C, device=dev)
def step():
    xg = x.detach().requires_grad_(True)
    for p in head.parameters(): p.grad = None
    loss, _ = head.forward_loss(xg, labels); loss.backward()
    return loss
b200.set_feature_pack_cache(0)
for on in (5, 1, 5, 1):
    _lib.gemm_set_sharing(on)
    for _ in range(3): step()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): l = step()
    e1.record(); torch.cuda.synchronize()
    prof = _lib.profile_read(); _lib.profile_enable(False)
    print("sharing", on, "loss", round(l.item(), 6), "step_ms(with events)", round(e0.elapsed_time(e1) / 10, 4),
          {k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items() if "gemm" in k})
