"""Small pass over K7 (test-time augmentation), K8 (optimizer steps) and the discriminator stack with the folded bias gradients
for compute-sanitizer (memcheck): ragged sizes, every template variant, unaligned tensor tails."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
for C, shapes, flips, divs, H, W in [(19, [(9, 17), (9, 17)], [False, True], (2,), 65, 129), (2, [(11, 11)], [True], (), 88, 88),
                                     (7, [(8, 16), (5, 9), (12, 20)], [False, True, False], (3,), 64, 131),
                                     (30, [(8, 8)] * 8, [False, True] * 4, (4, 2), 33, 257), (20, [(3, 3)], [False], (), 1, 1)]:
    members = [torch.randn(1, C, h, w, device=dev, generator=g) for h, w in shapes]
    lab = torch.randint(0, C, (1, H, W), device=dev, generator=g)
    lab[torch.rand(1, H, W, device=dev, generator=g) < 0.1] = 255
    for exact in (False, True):
        cm, pred, probs = _lib.tta_argmax_confusion(members, flips, (H, W), labels=lab, divisors=divs, want_pred=True, want_probs=True,
                                                    div_exact=exact)
    torch.cuda.synchronize()
    print("k7 ok", C, len(members), H, W, int(cm.sum()), float(probs.sum()) / (H * W))

shapes = [(19, 64, 3, 3), (19,), (4097,), (3, 5), (1,)] * 5            # 25 tensors: two launches, odd lengths, unaligned views
base = torch.randn(sum(torch.Size(s).numel() for s in shapes) + 64, device=dev, generator=g)
ps, off = [], 1                                                        # start one element in: 4-byte aligned only
for s in shapes:
    nel = torch.Size(s).numel()
    ps.append(torch.nn.Parameter(base[off:off + nel].view(s)))
    off += nel
for cls, kw in ((b200.FusedSGD, dict(lr=0.1, momentum=0.9, weight_decay=1e-3, nesterov=True)), (b200.FusedSGD, dict(lr=0.1)),
                (b200.FusedAdam, dict(lr=1e-3, weight_decay=1e-2))):
    opt = cls(ps, **kw)
    for _ in range(3):
        for p in ps:
            p.grad = torch.randn(p.shape, device=dev, generator=g)
        opt.step()
    torch.cuda.synchronize()
    print("k8 ok", cls.__name__, kw, float(base.abs().max()))

for n, cin, ndf, C, h, w in [(2, 64, 32, 19, 9, 11), (1, 24, 16, 3, 20, 27), (3, 128, 64, 2, 33, 17)]:
    D = b200.PixelDiscriminator(cin, ndf, num_classes=C).to(dev)
    x = torch.relu(torch.randn(n, cin, h, w, device=dev, generator=g)).requires_grad_(True)
    D(x).backward(torch.randn(n, 2 * C, h, w, device=dev, generator=g))
    torch.cuda.synchronize()
    print("k6 ok", n, cin, ndf, C, h, w, float(x.grad.abs().max()), [float(p.grad.abs().max()) for p in D.parameters()][1::2])
