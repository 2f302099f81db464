"""Short program for ncu: 2 warm-up + 1 profiled FADA iteration after the backbone (bench.py's adv_step, aspp_fada.py:91-125).
    python profiles/prof_adv.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth

RATES = [6, 12, 18, 24]
an, cin, ah, aw, aH, aW, aC = synth.WORKLOADS["deeplabv2_r101_adv"]
dev = torch.device("cuda", 0)
torch.manual_seed(4321)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, aC).to(dev)
model_D = b200.PixelDiscriminator(cin, 256, num_classes=aC).to(dev)
src = synth.make_features(an, cin, ah, aw, seed=555, device=dev)
tgt = synth.make_features(an, cin, ah, aw, seed=777, device=dev)
lab = synth.make_labels(an, aH, aW, aC, seed=555, device=dev)
size = (aH, aW)
b200.set_feature_pack_cache(2)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    b200.clear_feature_pack_cache()
    for p in list(head.parameters()) + list(model_D.parameters()):
        p.grad = None
    src_fea = src.detach().requires_grad_(True)
    tgt_fea = tgt.detach().requires_grad_(True)
    loss_seg, src_lr = head.forward_loss(src_fea, lab, temperature=1.8)
    loss_seg.backward()
    with torch.no_grad():
        tgt_lr = head.logits(tgt_fea)
    loss_adv = 0.001 * model_D.forward_soft_loss(tgt_fea, tgt_lr, size, slot=0)
    loss_adv.backward()
    for p in model_D.parameters():
        p.grad = None
    loss_d_src = 0.5 * model_D.forward_soft_loss(src_fea.detach(), src_lr, size, slot=0)
    loss_d_src.backward()
    loss_d_tgt = 0.5 * model_D.forward_soft_loss(tgt_fea.detach(), tgt_lr, size, slot=1)
    loss_d_tgt.backward()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", loss_seg.item(), loss_adv.item(), loss_d_src.item(), loss_d_tgt.item())
