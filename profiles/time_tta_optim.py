"""Timing + exactness diagnostics of K7 (fused test-time augmentation) and K8 (fused optimizer steps), and of the discriminator
backward with the bias gradients folded into the dgrad epilogues.  python profiles/time_tta_optim.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib
from oracle import torch_oracle as to

out = {}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
C, H, W = 19, 1024, 2048
labels = torch.randint(0, C, (1, H, W), device="cuda")
labels[torch.rand(1, H, W, device="cuda") < 0.1] = 255
for name, shapes, flips, divs in (("flip", [(128, 256)] * 2, [False, True], (2,)),
                                  ("flip_64x128", [(64, 128)] * 2, [False, True], (2,)),
                                  ("ms3_flip", [(90, 179), (90, 179), (128, 256), (128, 256), (167, 333), (167, 333)], [False, True] * 3, (3, 2))):
    members = [torch.randn(1, C, h, w, device="cuda") for h, w in shapes]
    want = to.tta_probabilities(members, flips, (H, W), divs)
    rec = {}
    for mode in (False, True):
        _, pred, probs = _lib.tta_argmax_confusion(members, flips, (H, W), divisors=divs, want_pred=True, want_probs=True, div_exact=mode)
        rec["div_exact" if mode else "div_recip"] = dict(probs_bit_equal=bool(torch.equal(probs.unsqueeze(0), want)),
                                                         max_abs=float((probs.unsqueeze(0) - want).abs().max()),
                                                         pred_mismatch=int((pred.unsqueeze(0) != want.max(1)[1]).sum()))
    cm = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    rec["fused_pred_cm_ms"] = timeit(lambda: _lib.tta_argmax_confusion(members, flips, (H, W), labels=labels, divisors=divs, cm=cm, want_pred=True))
    if len(members) <= 2:
        _lib.tta_set_row_walk(False)
        rec["fused_pred_cm_per_pixel_kernel_ms"] = timeit(lambda: _lib.tta_argmax_confusion(members, flips, (H, W), labels=labels, divisors=divs, cm=cm, want_pred=True))
        _lib.tta_set_row_walk(True)
    rec["fused_cm_only_ms"] = timeit(lambda: _lib.tta_argmax_confusion(members, flips, (H, W), labels=labels, divisors=divs, cm=cm))
    rec["fused_probs_ms"] = timeit(lambda: _lib.tta_argmax_confusion(members, flips, (H, W), divisors=divs, want_probs=True))

    def torch_path():
        p = to.tta_probabilities(members, flips, (H, W), divs)
        pd = p.max(1)[1]
        return torch.bincount((labels.flatten().clamp(max=C) * C + pd.flatten())[labels.flatten() != 255], minlength=C * C)
    rec["torch_cuda_ops_ms"] = timeit(torch_path, iters=5)
    rec["algorithmic_MB"] = (8 * H * W + sum(4 * C * h * w for h, w in shapes)) / 1e6
    out["k7_" + name] = rec
    print(name, json.dumps(rec))

# ---- K8
for name, shapes, kind, hyper in (("head_sgd", [(19, 2048, 3, 3), (19,)] * 4, "sgd", dict(lr=0.0025, momentum=0.9, weight_decay=5e-4)),
                                  ("disc_adam", [(256, 2048, 3, 3), (256,), (128, 256, 3, 3), (128,), (19, 128, 3, 3), (19,), (19, 128, 3, 3), (19,)],
                                   "adam", dict(lr=1e-4, betas=(0.9, 0.99)))):
    rec = {}
    n_el = 0
    for label, cls, kw in (("fused", b200.FusedSGD if kind == "sgd" else b200.FusedAdam, {}),
                           ("torch_foreach", torch.optim.SGD if kind == "sgd" else torch.optim.Adam, dict(foreach=True)),
                           ("torch_fused", torch.optim.SGD if kind == "sgd" else torch.optim.Adam, dict(fused=True))):
        ps = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.01) for s in shapes]
        n_el = sum(p.numel() for p in ps)
        for p in ps:
            p.grad = torch.randn_like(p) * 0.01
        try:
            opt = cls(ps, **hyper, **kw)
            rec[label + "_us"] = 1e3 * timeit(opt.step, iters=50, warm=5)
        except Exception as e:  # noqa: BLE001
            rec[label + "_us"] = repr(e)
    bytes_per = 20 if kind == "sgd" else 28
    rec["elements"] = n_el
    rec["algorithmic_MB"] = n_el * bytes_per / 1e6
    if isinstance(rec["fused_us"], float):
        rec["fused_GBps"] = n_el * bytes_per / rec["fused_us"] / 1e3
    out["k8_" + name] = rec
    print(name, json.dumps(rec))

# ---- discriminator fwd + bwd at the adversarial config (bias gradients folded into the dgrad epilogues)
N, h, w = 4, 64, 128
D = b200.PixelDiscriminator(2048, 256, num_classes=C).cuda()
x = torch.relu(torch.randn(N, 2048, h, w, device="cuda"))
go = torch.randn(N, 2 * C, h, w, device="cuda") * 1e-3


def fwdbwd(need_x):
    inp = x.detach().requires_grad_(need_x)
    for p in D.parameters():
        p.grad = None
    D(inp).backward(go)


out["disc_fwd_bwd_ms"] = dict(dx_dw=timeit(lambda: fwdbwd(True), iters=10), dw_only=timeit(lambda: fwdbwd(False), iters=10))
print("disc", json.dumps(out["disc_fwd_bwd_ms"]))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/time_tta_optim.json", "w") as f:
    json.dump(out, f, indent=1)
