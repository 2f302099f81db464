"""Small pass over every kernel family for compute-sanitizer (memcheck): ragged sizes, both label paths, eval + train."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib, ops

RATES = [6, 12, 18, 24]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
for (n, cin, C, h, w, H, W) in [(2, 64, 19, 9, 17, 65, 129), (1, 64, 2, 11, 11, 88, 88), (1, 128, 7, 8, 16, 64, 128), (1, 64, 30, 8, 8, 64, 64)]:
    head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
    x = torch.relu(torch.randn(n, cin, h, w, device=dev, generator=g)).requires_grad_(True)
    lab = torch.randint(0, C, (n, H, W), device=dev, generator=g)
    lab[torch.rand(n, H, W, device=dev, generator=g) < 0.1] = 255
    loss, lg = head.forward_loss(x, lab)
    loss.backward()
    out = head(x.detach(), (H, W))
    cm, pred = _lib.upsample_argmax_confusion(lg, lab, (H, W), want_pred=True)
    if cin % 8 == 0:
        xb = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        l2, _ = head.forward_loss(xb, lab)
        l2.backward()
    x2 = lg.detach().clone().requires_grad_(True)
    l3 = ops.upsample_cross_entropy(x2, lab)
    l3.backward()
    torch.cuda.synchronize()
    print("ok", n, cin, C, h, w, H, W, float(loss), int(cm.sum()))
