import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib, synth
RATES = [6, 12, 18, 24]
n, cin, h, w, H, W, C = synth.WORKLOADS["train_b8_512x1024"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
labels = synth.make_labels(n, H, W, C, device=dev)
def step():
    xg = x.detach().requires_grad_(True)
    for p in head.parameters(): p.grad = None
    loss, _ = head.forward_loss(xg, labels); loss.backward()
    return loss
for on in (1, 3, 1, 3):
    _lib.gemm_set_sharing(on)
    for _ in range(3): step()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): l = step()
    e1.record(); torch.cuda.synchronize()
    prof = _lib.profile_read(); _lib.profile_enable(False)
    print("seam sharing", on, "loss", round(l.item(), 6), "step_ms", round(e0.elapsed_time(e1) / 10, 4), {k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items() if "gemm" in k})
