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
