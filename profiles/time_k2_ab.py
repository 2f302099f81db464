"""K2 A/B: CTA-tile kernel (variant 0) vs warp-tile kernel (variant 1), int64 and uint8 labels, at the bench shape (and others):
CUDA-event time per launch of the main kernel, loss / low-res gradient agreement between the two, run-to-run determinism.
    python profiles/time_k2_ab.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rnd_semantic_segmentation_b200 import _lib, ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
for (n, C, h, w, H, W, T) in [(8, 19, 64, 128, 512, 1024, 1.0), (2, 19, 65, 129, 512, 1024, 1.8), (16, 2, 44, 44, 352, 352, 1.0), (1, 19, 64, 128, 512, 1024, 1.0)]:
    sets = []
    for i in range(4):
        lg = 2.0 * torch.randn(n, C, h, w, device=dev, generator=g)
        lab = torch.randint(0, C, (n, H, W), device=dev, generator=g)
        blocks = torch.rand(n, H // 64 + 1, W // 64 + 1, device=dev, generator=g) < 0.1
        lab[blocks.repeat_interleave(64, 1).repeat_interleave(64, 2)[:, :H, :W]] = 255
        sets.append((lg, lab, lab.to(torch.uint8)))
    res = {}
    for variant in (0, 1):
        _lib.upsample_ce_set_variant(2 * variant)          # 2: the warp-tile kernel wherever it is eligible
        for u8 in (False, True):
            _lib.profile_enable(True)
            outs = []
            for rep in range(5):
                for lg, lab, lab8 in sets:
                    x = lg.clone().requires_grad_(True)
                    loss = ops.upsample_cross_entropy(x, lab8 if u8 else lab, 255, T)
                    loss.backward()
                    if rep == 0:
                        outs.append((loss.detach().clone(), x.grad.clone()))
            torch.cuda.synchronize()
            ms, cnt = _lib.profile_read()["upsample_ce_main"]
            _lib.profile_enable(False)
            # determinism
            x = sets[0][0].clone().requires_grad_(True)
            l2 = ops.upsample_cross_entropy(x, sets[0][2] if u8 else sets[0][1], 255, T)
            l2.backward()
            det = bool(torch.equal(l2.detach(), outs[0][0]) and torch.equal(x.grad, outs[0][1]))
            res[(variant, u8)] = outs
            print(f"shape N={n} C={C} {h}x{w}->{H}x{W} T={T} variant={variant} u8={u8}: {ms / cnt * 1e3:7.2f} us/launch  loss={outs[0][0].item():.6f} deterministic={det}")
    for u8 in (False, True):
        dl = max(abs(a[0].item() - b[0].item()) / abs(b[0].item()) for a, b in zip(res[(1, u8)], res[(0, u8)]))
        dg = max(((a[1] - b[1]).abs().max() / b[1].abs().max()).item() for a, b in zip(res[(1, u8)], res[(0, u8)]))
        print(f"   v1 vs v0 (u8={u8}): max rel loss diff {dl:.2e}, max rel grad diff {dg:.2e}")
    same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(res[(1, False)], res[(1, True)]))
    print("   v1 int64 == v1 uint8 bit-identical:", same)
_lib.upsample_ce_set_variant(1)
