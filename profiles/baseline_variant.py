"""The north_star's prescribed baseline for the head -- four dilated 3x3 implicit GEMMs (halo tiles fetched by TMA, N = 19 classes
padded to the MMA width) summed -- measured beside the tap-packed GEMM head that is the product ("keep that as the
baseline variant and measure both").  The implicit-GEMM kernel is K6's (csrc/conv_sm100.cu, `conv3x3_forward(..., dilation=r)`,
fp32 NCHW output); the four results are added with three elementwise adds, as classifier.py:27-29 does.
    python profiles/baseline_variant.py            (CUDA-event timing + agreement)
    ncu --set full -k regex:conv_gemm_kernel ...   (tensor-pipe % of the baseline's kernel)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import _lib, synth

RATES = [6, 12, 18, 24]
dev = torch.device("cuda", 0)
for name in ("train_b8_512x1024", "eval_1024x2048", "deeplabv2_r101_src"):
    n, cin, h, w, H, W, C = synth.WORKLOADS[name]
    torch.manual_seed(0)
    head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev).eval()
    x = synth.make_features(n, cin, h, w, seed=3, device=dev)
    Xp = _lib.aspp_pack_features(x).view(n, h, w, cin)                       # bf16 NHWC: the operand layout of both variants
    packs = [_lib.conv3x3_pack_weights([m.weight.detach()])[0] for m in head.conv2d_list]
    biases = [m.bias.detach() for m in head.conv2d_list]

    def baseline():
        out = _lib.conv3x3_forward(Xp, packs[0], biases[0], None, dilation=RATES[0], out_f32_nchw=True)
        for r in range(1, 4):
            out += _lib.conv3x3_forward(Xp, packs[r], biases[r], None, dilation=RATES[r], out_f32_nchw=True)
        return out

    def product():
        with torch.no_grad():
            return head.logits(Xp.permute(0, 3, 1, 2))                       # zero-copy bf16 channels_last operand: GEMM + gather only

    a, b = baseline(), product()
    err = ((a - b).abs().max() / b.abs().max()).item()
    res = {}
    for fn, key in ((baseline, "four dilated implicit GEMMs, N = 19 padded (baseline)"), (product, "tap-packed GEMM + gather (product)")):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[key] = e0.elapsed_time(e1) / 10
    flops = 2.0 * 36 * cin * C * n * h * w
    print(f"{name}: max rel difference {err:.2e}")
    for key, ms in res.items():
        print(f"   {key}: {ms * 1e3:8.1f} us   {flops / (ms * 1e-3) / 1e12:7.1f} TFLOP/s algorithmic (dense 36 taps, C = {C})")
