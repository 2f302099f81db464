import numpy as np
import torch

RATES = [6, 12, 18, 24]


def rel_err(a, b):
    """max-abs difference over max-abs reference (the parity metric of BASELINE.md)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def effective_bf16_head(head_oracle):
    """Oracle head fed the SAME bf16-rounded packed operands the CUDA head uses: every off-centre tap weight
    rounded to bf16; the four centre taps pre-summed in fp32, rounded once, carried by branch 0."""
    import copy
    eff = copy.deepcopy(head_oracle)
    with torch.no_grad():
        centre = sum(m.weight[:, :, 1, 1].float() for m in head_oracle.conv2d_list)
        for i, m in enumerate(eff.conv2d_list):
            m.weight.copy_(bf16_round(m.weight))
            m.weight[:, :, 1, 1] = bf16_round(centre) if i == 0 else 0.0
    return eff


def make_labels(n, H, W, C, p_ignore, seed):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, C, (n, H, W), generator=g)
    lab[torch.rand(n, H, W, generator=g) < p_ignore] = 255
    return lab


# ---- PixelDiscriminator stack: an oracle applying exactly the bf16 roundings of the CUDA stack (fp64 math otherwise) ----
class _RoundFwd(torch.autograd.Function):
    """bf16-round the value, pass the gradient through (a rounded GEMM operand)."""

    @staticmethod
    def forward(ctx, t):
        return bf16_round(t.float()).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """identity forward, bf16-round the gradient (the gradient tensors the CUDA stack stores as bf16 NHWC)."""

    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return bf16_round(g.float()).to(g.dtype)


class _LeakyReLUGivenMask(torch.autograd.Function):
    """LeakyReLU whose backward uses a GIVEN sign mask (the CUDA stack derives it from its own stored activation, so an
    fp32-vs-fp64 sign flip of a pre-activation within rounding distance of 0 cannot make the two sides disagree)."""

    @staticmethod
    def forward(ctx, z, mask_pos, slope):
        ctx.save_for_backward(mask_pos)
        ctx.slope = slope
        return torch.where(z > 0, z, z * slope)

    @staticmethod
    def backward(ctx, g):
        mask_pos, = ctx.saved_tensors
        return g * torch.where(mask_pos, 1.0, ctx.slope).to(g.dtype), None, None


def discriminator_bf16_oracle(D, x, masks=None, slope=0.2):
    """D: an fp64 module with the reference's layout (D.0, D.2, cls1, cls2); x fp64 (requires_grad as needed).
    masks: optional (A1 > 0, A2 > 0) boolean NCHW tensors taken from the CUDA stack's stored activations."""
    import torch.nn.functional as F
    r = _RoundFwd.apply
    z1 = _RoundBwd.apply(F.conv2d(r(x), r(D.D[0].weight), D.D[0].bias, padding=1))
    a1 = r(_LeakyReLUGivenMask.apply(z1, (z1 > 0) if masks is None else masks[0], slope))
    z2 = _RoundBwd.apply(F.conv2d(a1, r(D.D[2].weight), D.D[2].bias, padding=1))
    a2 = r(_LeakyReLUGivenMask.apply(z2, (z2 > 0) if masks is None else masks[1], slope))
    out = torch.cat((F.conv2d(a2, r(D.cls1.weight), D.cls1.bias, padding=1), F.conv2d(a2, r(D.cls2.weight), D.cls2.bias, padding=1)), 1)
    return _RoundBwd.apply(out)
