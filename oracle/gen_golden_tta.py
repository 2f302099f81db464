"""Generate tests/golden/tta.npz and tests/golden/optim.npz by running the REFERENCE's own code (imported by path).

Run in the authoring container only (needs /root/reference):

    python -m oracle.gen_golden_tta

tta.npz: ``inference(..., flip=True)`` and ``multi_scale_inference(...)`` of core/utils/utility.py:179-209 driven with the
reference's ``ASPP_Classifier_V2`` behind a small strided-conv backbone (so that mirroring / rescaling the image changes the
low-res logits), together with the low-res logits of every ensemble member in the reference's order -- the inputs of the
fused kernel.  optim.npz: five iterations of the optimizers exactly as core/trainers/aspp_trainer.py:25-26,77-81,94-95 and
core/adapters/fada_adapter.py:24 build and drive them (poly schedule from the reference's adjust_learning_rate).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
RATES = [6, 12, 18, 24]
SCALES = [0.7, 1.0, 1.3]


class Recorder(torch.nn.Module):
    """Wraps the head and keeps every low-res output it produced, in call order."""

    def __init__(self, head):
        super().__init__()
        self.head = head
        self.outputs = []

    def forward(self, x, size=None):
        out = self.head(x, size)
        self.outputs.append(out.detach().clone())
        return out


def gen_tta(ref):
    torch.manual_seed(51)
    gen = torch.Generator().manual_seed(52)
    ncls, cin = 19, 12
    backbone = torch.nn.Sequential(torch.nn.Conv2d(3, cin, 3, stride=4, padding=1), torch.nn.ReLU())
    head = ref.ASPP_Classifier_V2(cin, RATES, RATES, ncls)
    with torch.no_grad():
        for m in head.conv2d_list:
            m.weight.mul_(40.0)                      # logits sigma ~ 1
    image = torch.randn(1, 3, 48, 64, generator=gen)
    label = torch.randint(0, ncls, (1, 50, 70), generator=gen)
    label[torch.rand(1, 50, 70, generator=gen) < 0.1] = 255
    arrs = dict(image=image.numpy(), label=label.numpy(), num_classes=np.int64(ncls), scales=np.asarray(SCALES))

    rec = Recorder(head)
    probs = ref.inference(backbone, rec, image, label, flip=True)                       # utility.py:179-191
    both = rec.outputs[0]                                                               # [2,C,h,w]: plain, mirrored
    arrs.update({"flip.member0": both[0:1].numpy(), "flip.member1": both[1:2].numpy(), "flip.probs": probs.numpy(),
                 "flip.pred": probs.max(1)[1].numpy()})

    for name, flip in (("ms", True), ("ms_noflip", False)):
        rec = Recorder(head)
        probs = ref.multi_scale_inference(backbone, rec, image, label, flip=flip, scales=SCALES)   # utility.py:193-209
        for k, o in enumerate(rec.outputs):
            arrs[f"{name}.member{k}"] = o.numpy()
        arrs[f"{name}.n_members"] = np.int64(len(rec.outputs))
        arrs[f"{name}.probs"] = probs.numpy()
        arrs[f"{name}.pred"] = probs.max(1)[1].numpy()
    np.savez_compressed(os.path.join(OUT, "tta.npz"), **arrs)
    print("tta: members", [tuple(arrs[f"ms.member{k}"].shape) for k in range(6)])


def gen_optim(ref):
    gen = torch.Generator().manual_seed(61)
    shapes = [(19, 24, 3, 3), (19,), (5, 7), (4099,)]            # conv weight, bias, an odd matrix, one > 4096-element vector
    p0 = [torch.randn(s, generator=gen) * 0.05 for s in shapes]
    steps = 5
    grads = [[torch.randn(s, generator=gen) * 0.01 for s in shapes] for _ in range(steps)]
    base_lr, max_iter, power = 2.5e-4, 20, 0.9
    lrs = [ref.adjust_learning_rate('poly', base_lr, it, max_iter, power) for it in range(steps)]   # aspp_trainer.py:77
    arrs = dict(steps=np.int64(steps), n=np.int64(len(shapes)), lrs=np.asarray(lrs, dtype=np.float64),
                base_lr=np.float64(base_lr), max_iter=np.int64(max_iter), power=np.float64(power))
    for i, p in enumerate(p0):
        arrs[f"p0.{i}"] = p.numpy()
        for k in range(steps):
            arrs[f"g{k}.{i}"] = grads[k][i].numpy()

    def run(opt_ctor, scale):
        ps = [torch.nn.Parameter(p.clone()) for p in p0]
        opt = opt_ctor(ps)
        for k in range(steps):
            for grp in opt.param_groups:                          # aspp_trainer.py:80-81
                grp['lr'] = lrs[k] * scale
            opt.zero_grad()
            for p, g in zip(ps, grads[k]):
                p.grad = g.clone()
            opt.step()
        return ps, opt

    # optimizer_cls: SGD, lr = BASE_LR * 10, momentum 0.9, weight decay 5e-4 (aspp_trainer.py:26, yaml SOLVER defaults)
    ps, opt = run(lambda ps: torch.optim.SGD(ps, lr=base_lr * 10, momentum=0.9, weight_decay=5e-4), 10)
    for i, p in enumerate(ps):
        arrs[f"sgd.p.{i}"] = p.detach().numpy()
        arrs[f"sgd.buf.{i}"] = opt.state[p]["momentum_buffer"].numpy()
    # optimizer_D: Adam, betas (0.9, 0.99) (fada_adapter.py:24)
    ps, opt = run(lambda ps: torch.optim.Adam(ps, lr=1e-4, betas=(0.9, 0.99)), 0.4)
    for i, p in enumerate(ps):
        arrs[f"adam.p.{i}"] = p.detach().numpy()
        arrs[f"adam.m.{i}"] = opt.state[p]["exp_avg"].numpy()
        arrs[f"adam.v.{i}"] = opt.state[p]["exp_avg_sq"].numpy()
    np.savez_compressed(os.path.join(OUT, "optim.npz"), **arrs)
    print("optim: lrs", lrs)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gen_tta(ref)
    gen_optim(ref)


if __name__ == "__main__":
    main()
