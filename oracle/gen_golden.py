"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported by path).

Run in the authoring container only (needs /root/reference):

    python -m oracle.gen_golden

The reference ships no golden vectors for this path, so
these fixtures -- produced by the reference's own ``ASPP_Classifier_V2``,
``PixelDiscriminator``, ``soft_label_cross_entropy``, ``inference``,
``confusion_matrix``, ``intersectionAndUnion`` and ``AverageMeter`` under torch
2.11 CPU -- are what pins the oracle.  Shapes are small so the files stay tiny.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
RATES = [6, 12, 18, 24]


def _labels(gen, shape, num_classes, p_ignore):
    lab = torch.randint(0, num_classes, shape, generator=gen)
    lab[torch.rand(shape, generator=gen) < p_ignore] = 255
    return lab


def _sd(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def gen_head(ref, name, seed, cin, ncls, n, h, w, H, W, p_ignore=0.1, temperature=1.0):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    head = ref.ASPP_Classifier_V2(cin, RATES, RATES, ncls)
    x = torch.relu(torch.randn(n, cin, h, w, generator=gen)).requires_grad_(True)
    labels = _labels(gen, (n, H, W), ncls, p_ignore)
    out_lr = head(x)
    out_hr = head(x, (H, W))
    scaled = out_hr.div(temperature) if temperature != 1.0 else out_hr
    loss = torch.nn.CrossEntropyLoss(ignore_index=255)(scaled, labels)   # as aspp_trainer.py:61,91
    loss.backward()
    arrs = dict(x=x.detach().numpy(), labels=labels.numpy(), logits_lr=out_lr.detach().numpy(),
                logits_hr=out_hr.detach().numpy(), loss=np.float32(loss.item()),
                grad_x=x.grad.numpy(), temperature=np.float32(temperature),
                num_classes=np.int64(ncls))
    for k, v in _sd(head).items():
        arrs["sd." + k] = v
    for k, p in head.named_parameters():
        arrs["grad." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(name, "loss", loss.item())


def gen_all_ignored(ref):
    torch.manual_seed(7)
    head = ref.ASPP_Classifier_V2(8, RATES, RATES, 5)
    x = torch.relu(torch.randn(1, 8, 7, 7))
    labels = torch.full((1, 20, 20), 255, dtype=torch.int64)
    loss = torch.nn.CrossEntropyLoss(ignore_index=255)(head(x, (20, 20)), labels)
    arrs = dict(x=x.numpy(), labels=labels.numpy(), loss=np.float32(loss.item()), num_classes=np.int64(5))
    for k, v in _sd(head).items():
        arrs["sd." + k] = v
    np.savez_compressed(os.path.join(OUT, "head_all_ignored.npz"), **arrs)
    print("all ignored loss", loss.item())


def gen_soft_ce(ref):
    gen = torch.Generator().manual_seed(11)
    pred = (3 * torch.randn(2, 6, 7, 9, generator=gen)).requires_grad_(True)
    soft = torch.softmax(2 * torch.randn(2, 6, 7, 9, generator=gen), dim=1)
    soft[:, 3:] = 0                      # the [soft, 0] slot pattern of aspp_fada.py:111
    wts = torch.rand(2, 7, 9, generator=gen)
    loss = ref.soft_label_cross_entropy(pred, soft)
    loss.backward()
    g0 = pred.grad.clone()
    pred.grad = None
    loss_w = ref.soft_label_cross_entropy(pred, soft, wts)
    loss_w.backward()
    np.savez_compressed(os.path.join(OUT, "soft_ce.npz"), pred=pred.detach().numpy(), soft=soft.numpy(),
                        weights=wts.numpy(), loss=np.float32(loss.item()), grad=g0.numpy(),
                        loss_w=np.float32(loss_w.item()), grad_w=pred.grad.numpy())
    print("soft_ce", loss.item(), loss_w.item())


def gen_discriminator(ref):
    torch.manual_seed(21)
    gen = torch.Generator().manual_seed(22)
    ncls = 3
    D = ref.PixelDiscriminator(24, 16, num_classes=ncls)
    x = torch.relu(torch.randn(2, 24, 9, 11, generator=gen))
    out_lr = D(x)
    out_hr = D(x, (20, 27))
    # the adversarial-stage inline sequence, aspp_fada.py:93-94,99-100,110-111,124
    seg_hr = 4 * torch.randn(2, ncls, 20, 27, generator=gen)
    soft = F.softmax(seg_hr.div(1.8), dim=1).detach()
    soft[soft > 0.9] = 0.9
    q0 = torch.cat((soft, torch.zeros_like(soft)), dim=1)
    q1 = torch.cat((torch.zeros_like(soft), soft), dim=1)
    l0 = ref.soft_label_cross_entropy(out_hr, q0)
    l1 = ref.soft_label_cross_entropy(out_hr, q1)
    g0 = torch.autograd.grad(l0, out_lr if False else D.cls1.weight, retain_graph=True)[0]
    arrs = dict(x=x.numpy(), out_lr=out_lr.detach().numpy(), out_hr=out_hr.detach().numpy(),
                seg_hr=seg_hr.numpy(), soft=soft.numpy(), loss_slot0=np.float32(l0.item()),
                loss_slot1=np.float32(l1.item()), grad_cls1_weight_slot0=g0.numpy(),
                num_classes=np.int64(ncls))
    for k, v in _sd(D).items():
        arrs["sd." + k] = v
    np.savez_compressed(os.path.join(OUT, "discriminator.npz"), **arrs)
    print("discriminator", l0.item(), l1.item())


def gen_eval(ref):
    torch.manual_seed(31)
    gen = torch.Generator().manual_seed(32)
    ncls, cin = 19, 16
    head = ref.ASPP_Classifier_V2(cin, RATES, RATES, ncls)
    with torch.no_grad():
        for m in head.conv2d_list:
            m.weight.mul_(30.0)          # logits sigma ~ 1 so argmax is well separated
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(NUM_CLASSES=ncls))
    meter = ref.AverageMeter()
    frames = []
    cmt = torch.zeros(ncls, ncls, dtype=torch.int64)
    for f in range(3):
        x = torch.relu(torch.randn(1, cin, 8, 11, generator=gen))
        y = _labels(gen, (1, 24, 31), ncls, 0.15)
        probs = ref.inference(torch.nn.Identity(), head, x, y, flip=False)   # utility.py:179-191
        pred = probs.max(1)[1]                                                 # aspp_tester.py:63
        cm = ref.confusion_matrix(cfg, torch.flatten(pred), torch.flatten(y)) # :67-69
        cmt = cmt + cm
        i, u, t, r = ref.intersectionAndUnion(pred.numpy(), y.numpy(), ncls, 255)  # numpy twin :133-145
        i, u, t, r = (a.astype(np.float32) for a in (i, u, t, r))               # GPU variant returns float32
        meter.update(i, u, t, r)
        with torch.no_grad():
            frames.append(dict(x=x.numpy(), y=y.numpy(), logits_lr=head(x).numpy(), probs=probs.numpy(),
                               pred=pred.numpy(), cm=cm.numpy(), I=i, U=u, T=t, R=r))
    arrs = {f"f{k}.{n}": v for k, fr in enumerate(frames) for n, v in fr.items()}
    arrs.update(cmt=cmt.numpy(), num_classes=np.int64(ncls),
                meter_iou_sum=np.asarray(meter.iou_sum), meter_f1_sum=np.asarray(meter.f1_sum),
                meter_intersection_sum=np.asarray(meter.intersection_sum),
                meter_union_sum=np.asarray(meter.union_sum), meter_target_sum=np.asarray(meter.target_sum),
                meter_res_sum=np.asarray(meter.res_sum), meter_count=np.int64(meter.count))
    for k, v in _sd(head).items():
        arrs["sd." + k] = v
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **arrs)
    print("eval cmt trace", int(cmt.trace()), "total", int(cmt.sum()))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gen_head(ref, "head_c19", seed=1, cin=24, ncls=19, n=2, h=13, w=17, H=40, W=53)
    gen_head(ref, "head_c2", seed=2, cin=16, ncls=2, n=3, h=9, w=9, H=30, W=30, p_ignore=0.05)
    gen_head(ref, "head_c19_T18", seed=3, cin=32, ncls=19, n=1, h=27, w=31, H=64, W=96, temperature=1.8)
    gen_all_ignored(ref)
    gen_soft_ce(ref)
    gen_discriminator(ref)
    gen_eval(ref)


if __name__ == "__main__":
    main()
