"""CPU oracle: the reference's hot path restated with the torch calls it makes.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line (relative to the reference root) whose behaviour it
restates.  All functions run on whatever device their inputs live on, so the
same code is the "GPU oracle" (torch CUDA ops) for the bit-exact argmax /
confusion-matrix claims and the CPU oracle for tolerance-based parity and for
the CPU baseline timing.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

ASPP_RATES = (6, 12, 18, 24)


# --------------------------------------------------------------------------
# a1/a2  ASPP_Classifier_V2  (core/models/classifiers/aspp/classifier.py:6-32)
# --------------------------------------------------------------------------
class AsppHeadOracle(nn.Module):
    """Four parallel dilated 3x3 convs whose outputs are summed.

    Follows classifier.py:7-24 (construction order, N(0, 0.01) weight init,
    default Conv2d bias init) and classifier.py:26-32 (sum order ((0+1)+2)+3,
    optional align_corners bilinear resize).  state_dict keys are
    ``conv2d_list.{i}.{weight,bias}`` exactly as in the reference.
    """

    def __init__(self, in_channels: int, dilation_series: Sequence[int],
                 padding_series: Sequence[int], num_classes: int):
        super().__init__()
        self.conv2d_list = nn.ModuleList(
            nn.Conv2d(in_channels, num_classes, 3, 1, padding=p, dilation=d, bias=True)
            for d, p in zip(dilation_series, padding_series))
        for conv in self.conv2d_list:                       # classifier.py:23-24
            conv.weight.data.normal_(0, 0.01)

    def forward(self, x, size=None):
        total = self.conv2d_list[0](x)                      # classifier.py:27
        for conv in list(self.conv2d_list)[1:]:             # classifier.py:28-29
            total = total + conv(x)
        if size is not None:                                # classifier.py:30-31
            total = upsample_bilinear_ac(total, size)
        return total


def head_forward(x: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                 rates: Sequence[int] = ASPP_RATES, size=None) -> torch.Tensor:
    """Functional form of classifier.py:26-32 for explicit weights."""
    total = None
    for w, b, r in zip(weights, biases, rates):
        y = F.conv2d(x, w, b, stride=1, padding=r, dilation=r)
        total = y if total is None else total + y
    if size is not None:
        total = upsample_bilinear_ac(total, size)
    return total


# --------------------------------------------------------------------------
# a3  F.interpolate(bilinear, align_corners=True)
#     (classifier.py:31, discriminator.py:49, utility.py:185)
# --------------------------------------------------------------------------
def upsample_bilinear_ac(x: torch.Tensor, size) -> torch.Tensor:
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=True)


# --------------------------------------------------------------------------
# a4  CrossEntropyLoss(ignore_index=255)
#     (aspp_trainer.py:61,91; aspp_fada.py:46,95; train_distill.py:112,136)
# --------------------------------------------------------------------------
def hard_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = 255) -> torch.Tensor:
    """mean over non-ignored pixels of -log_softmax(logits)[label]; NaN when none valid."""
    return F.cross_entropy(logits, labels, ignore_index=ignore_index)


# --------------------------------------------------------------------------
# a5  soft_label_cross_entropy  (core/utils/utility.py:172-177)
# --------------------------------------------------------------------------
def soft_label_cross_entropy(pred: torch.Tensor, soft_label: torch.Tensor,
                             pixel_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean over N*H*W of [w *] sum_c( -q_c * log_softmax(p)_c )."""
    per_px = torch.sum(-soft_label.float() * F.log_softmax(pred, dim=1), dim=1)
    if pixel_weights is not None:
        per_px = pixel_weights * per_px
    return per_px.mean()


# --------------------------------------------------------------------------
# a6  soft-label builder  (core/combos/aspp_fada.py:93-94,99-100,104-108,111,120,124)
# --------------------------------------------------------------------------
def build_soft_label(scaled_logits: torch.Tensor, slot: int, clamp: float = 0.9) -> torch.Tensor:
    """softmax over classes, values above ``clamp`` set to ``clamp``, detached, placed in
    channel slot 0 ([soft, 0]) or slot 1 ([0, soft]) of a 2C-channel tensor.

    ``scaled_logits`` are the full-resolution logits already divided by the
    temperature (aspp_fada.py:93-94 / :104)."""
    soft = F.softmax(scaled_logits, dim=1).detach()
    soft = torch.where(soft > clamp, torch.full_like(soft, clamp), soft)   # aspp_fada.py:100,108
    zeros = torch.zeros_like(soft)
    parts = (soft, zeros) if slot == 0 else (zeros, soft)                   # :111,:120 vs :124
    return torch.cat(parts, dim=1)


# --------------------------------------------------------------------------
# a7  PixelDiscriminator  (core/models/discriminator.py:31-50)
# --------------------------------------------------------------------------
class PixelDiscriminatorOracle(nn.Module):
    """conv3x3(in->ndf)+LReLU(0.2), conv3x3(ndf->ndf/2)+LReLU, two conv3x3 heads, cat."""

    def __init__(self, input_nc: int, ndf: int = 512, num_classes: int = 1):
        super().__init__()
        self.D = nn.Sequential(
            nn.Conv2d(input_nc, ndf, 3, 1, 1), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(ndf, ndf // 2, 3, 1, 1), nn.LeakyReLU(0.2, inplace=True))
        self.cls1 = nn.Conv2d(ndf // 2, num_classes, 3, 1, 1)
        self.cls2 = nn.Conv2d(ndf // 2, num_classes, 3, 1, 1)

    def forward(self, x, size=None):
        mid = self.D(x)
        both = torch.cat((self.cls1(mid), self.cls2(mid)), dim=1)          # discriminator.py:45-47
        if size is not None:
            both = upsample_bilinear_ac(both, size)                         # :48-49
        return both


# --------------------------------------------------------------------------
# a8/a9  inference(...) probabilities, then output.max(1)[1]
#        (core/utils/utility.py:179-191, core/testers/aspp_tester.py:63)
# --------------------------------------------------------------------------
def eval_probabilities(logits_lr: torch.Tensor, size) -> torch.Tensor:
    """flip=False branch of utility.py:179-191 starting from the head's low-res logits."""
    up = upsample_bilinear_ac(logits_lr, size)
    return F.softmax(up, dim=1)


def eval_argmax(logits_lr: torch.Tensor, size) -> torch.Tensor:
    """argmax over classes OF THE SOFTMAX PROBABILITIES (first index on ties), int64."""
    return eval_probabilities(logits_lr, size).max(1)[1]


# --------------------------------------------------------------------------
# 8f-3  test-time augmentation: inference(flip=True) and multi_scale_inference
#        (core/utils/utility.py:179-191 and :193-209)
# --------------------------------------------------------------------------
def inference(feature_extractor, classifier, image, label, flip: bool = True) -> torch.Tensor:
    """utility.py:179-191 line by line: [1,C,H,W] probabilities, flip-averaged when ``flip``."""
    size = label.shape[-2:]
    if flip:
        image = torch.cat([image, torch.flip(image, [3])], 0)            # :182
    with torch.no_grad():
        output = classifier(feature_extractor(image))                    # :184
    output = upsample_bilinear_ac(output, size)                          # :185
    output = F.softmax(output, dim=1)                                    # :186
    if flip:
        output = (output[0] + output[1].flip(2)) / 2                     # :188
    else:
        output = output[0]
    return output.unsqueeze(dim=0)


def multi_scale_inference(feature_extractor, classifier, image, label, flip: bool = True,
                          scales=(0.7, 1.0, 1.3)) -> torch.Tensor:
    """utility.py:193-209 line by line."""
    output = None
    size = image.shape[-2:]
    for s in scales:
        x = F.interpolate(image, size=(int(size[0] * s), int(size[1] * s)), mode='bilinear', align_corners=True)
        pred = inference(feature_extractor, classifier, x, label, flip=False)
        output = pred if output is None else output + pred
        if flip:
            x_flip = torch.flip(x, [3])
            pred = inference(feature_extractor, classifier, x_flip, label, flip=False)
            output = output + pred.flip(3)
    if flip:
        return output / len(scales) / 2
    return output / len(scales)


def tta_probabilities(members: Sequence[torch.Tensor], flips: Sequence[bool], size, divisors: Sequence[float] = ()) -> torch.Tensor:
    """The same ensembles starting from the members' low-res logits [1,C,h_m,w_m] (what the fused kernel is given):
    sum over members, in order, of the (un-mirrored) softmax of the upsampled logits, then the scalar divisions in order."""
    total = None
    for lg, fl in zip(members, flips):
        pr = F.softmax(upsample_bilinear_ac(lg, size), dim=1)
        if fl:
            pr = pr.flip(3)
        total = pr if total is None else total + pr
    for d in divisors:
        total = total / d
    return total


# --------------------------------------------------------------------------
# 8f-4  optimizer steps: torch.optim.SGD / Adam exactly as the reference builds them
#        (core/trainers/aspp_trainer.py:25-26,77-81,94-95; core/adapters/fada_adapter.py:24; core/utils/adapt_lr.py:12-17)
# --------------------------------------------------------------------------
def adjust_learning_rate(method, base_lr, iters, max_iter, power):
    """adapt_lr.py:12-17."""
    if method == 'poly':
        return base_lr * ((1 - float(iters) / max_iter) ** (power))
    raise NotImplementedError


def optimizer_steps(kind: str, params: Sequence[torch.Tensor], grads_per_step: Sequence[Sequence[torch.Tensor]],
                    lrs: Optional[Sequence[float]] = None, **hyper):
    """Run ``len(grads_per_step)`` steps of torch.optim.SGD / torch.optim.Adam (single-tensor eager path) on copies of
    ``params``; ``lrs[k]`` is written into the param group before step k as aspp_trainer.py:78-81 does.  Returns (params,
    state tensors per parameter: [momentum_buffer] for SGD, [exp_avg, exp_avg_sq] for Adam)."""
    ps = [torch.nn.Parameter(p.detach().clone()) for p in params]
    opt = (torch.optim.SGD if kind == "sgd" else torch.optim.Adam)(ps, foreach=False, **hyper)
    for k, grads in enumerate(grads_per_step):
        if lrs is not None:
            for grp in opt.param_groups:
                grp["lr"] = lrs[k]
        for p, g in zip(ps, grads):
            p.grad = g.detach().clone()
        opt.step()
    keys = ("momentum_buffer",) if kind == "sgd" else ("exp_avg", "exp_avg_sq")
    return [p.detach() for p in ps], [[opt.state[p].get(k) for k in keys] for p in ps]


# --------------------------------------------------------------------------
# a10  confusion_matrix  (core/utils/utility.py:347-359)
# --------------------------------------------------------------------------
def confusion_matrix_loop(num_classes: int, pd: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Literal restatement: one Python iteration per pixel, rows=truth, cols=prediction,
    truth == 255 skipped (hard-coded 255 in the reference).  Small inputs only."""
    cmt = torch.zeros(num_classes, num_classes, dtype=torch.int64)
    for t, p in zip(gt.tolist(), pd.tolist()):
        if t != 255:
            cmt[t, p] += 1
    return cmt


def confusion_matrix_bincount(num_classes: int, pd: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Vectorised equivalent of the loop above (equal for every input with valid indices)."""
    pd = pd.reshape(-1).to(torch.int64)
    gt = gt.reshape(-1).to(torch.int64)
    keep = gt != 255
    flat = gt[keep] * num_classes + pd[keep]
    return torch.bincount(flat, minlength=num_classes * num_classes).reshape(num_classes, num_classes).cpu()


# --------------------------------------------------------------------------
# a11  intersectionAndUnion[GPU]  (core/utils/utility.py:133-161)
# --------------------------------------------------------------------------
def intersection_and_union(pred: torch.Tensor, target: torch.Tensor, K: int, ignore_index: int = 255):
    """Returns float32 [K] x4 (intersection, union, target, output) as the reference does.

    The reference writes ``ignore_index`` into ``output`` where the target is ignored and
    then histograms values 0..K-1 (histc with bins=K,min=0,max=K-1 == bincount on small ints).
    """
    assert pred.dim() in (1, 2, 3)
    assert pred.shape == target.shape
    out = pred.reshape(-1).clone()
    tgt = target.reshape(-1)
    out[tgt == ignore_index] = ignore_index
    inter = out[out == tgt]

    def hist(v):
        v = v[(v >= 0) & (v <= K - 1)]
        return torch.bincount(v, minlength=K)[:K].to(torch.float32)

    area_i, area_o, area_t = hist(inter), hist(out), hist(tgt)
    return area_i, area_o + area_t - area_i, area_t, area_o


# --------------------------------------------------------------------------
# a12  AverageMeter  (core/utils/utility.py:24-72)
# --------------------------------------------------------------------------
class AverageMeterOracle:
    def __init__(self):
        self.reset()

    def reset(self):
        self.intersection_sum = 0
        self.union_sum = 0
        self.target_sum = 0
        self.res_sum = 0
        self.count = 0
        self.iou_sum = 0
        self.f1_sum = 0

    def update(self, intersection, union, target, res):
        self.iou_sum = self.iou_sum + intersection / (union + 1e-10)        # utility.py:44
        self.f1_sum = self.f1_sum + 2 * intersection / (target + res + 1e-10)  # :45
        self.intersection_sum = self.intersection_sum + intersection
        self.union_sum = self.union_sum + union
        self.target_sum = self.target_sum + target
        self.res_sum = self.res_sum + res
        self.count += 1

    def results(self):
        macro_iou = self.iou_sum / float(self.count)
        macro_f1 = self.f1_sum / float(self.count)
        micro_iou = self.intersection_sum / (self.union_sum + 1e-10)
        micro_f1 = 2 * self.intersection_sum / (self.target_sum + self.res_sum + 1e-10)
        return dict(macro_iou=macro_iou, macro_f1=macro_f1, micro_iou=micro_iou, micro_f1=micro_f1,
                    macro_mIoU=float(np.mean(macro_iou)), macro_mF1=float(np.mean(macro_f1)),
                    micro_mIoU=float(np.mean(micro_iou)), micro_mF1=float(np.mean(micro_f1)))


# --------------------------------------------------------------------------
# Whole steps (what the callers do around the hot path)
# --------------------------------------------------------------------------
def train_step_src(head: nn.Module, x: torch.Tensor, labels: torch.Tensor,
                   temperature: float = 1.0, ignore_index: int = 255):
    """aspp_trainer.py:88-92: out = head(x, size); loss = CE(out, y); loss.backward().

    Returns (loss, grad_x, [grad of each head parameter in state_dict order])."""
    x = x.detach().clone().requires_grad_(True)
    for p in head.parameters():
        p.grad = None
    out = head(x, labels.shape[-2:])
    if temperature != 1.0:
        out = out.div(temperature)                          # aspp_fada.py:93-94
    loss = hard_cross_entropy(out, labels, ignore_index)
    loss.backward()
    return loss.detach(), x.grad.detach(), [p.grad.detach().clone() for p in head.parameters()]


def eval_frame(head: nn.Module, x: torch.Tensor, labels: torch.Tensor, num_classes: int,
               ignore_index: int = 255, literal_loop: bool = False):
    """aspp_tester.py:57-74 for one frame, starting from layer4 features.

    Returns (pred int64 [1,H,W], cmt int64 [C,C], (I, U, T, R) float32 [C])."""
    with torch.no_grad():
        logits_lr = head(x)                                 # utility.py:183-184 (size=None)
    pred = eval_argmax(logits_lr, labels.shape[-2:])        # utility.py:185-186, aspp_tester.py:63
    cm_fn = confusion_matrix_loop if literal_loop else confusion_matrix_bincount
    cmt = cm_fn(num_classes, pred.flatten(), labels.flatten())
    iutr = intersection_and_union(pred, labels, num_classes, ignore_index)
    return pred, cmt, iutr


def fada_losses(head: nn.Module, model_D: nn.Module, src_fea, tgt_fea, src_label,
                temperature: float = 1.8, ignore_index: int = 255):
    """The post-backbone arithmetic of one adversarial iteration, aspp_fada.py:91-125,
    without optimizer steps (so the three D passes all see the same D weights).

    Returns dict of the five scalar losses the reference logs (:129-133)."""
    size = src_label.shape[-2:]
    src_pred = head(src_fea, size).div(temperature)
    loss_seg = hard_cross_entropy(src_pred, src_label, ignore_index)
    src_q0 = build_soft_label(src_pred, slot=0)
    tgt_pred = head(tgt_fea, size).div(temperature)
    tgt_q0 = build_soft_label(tgt_pred, slot=0)
    tgt_q1 = build_soft_label(tgt_pred, slot=1)
    loss_adv_tgt = 0.001 * soft_label_cross_entropy(model_D(tgt_fea, size), tgt_q0)
    loss_D_src = 0.5 * soft_label_cross_entropy(model_D(src_fea.detach(), size), src_q0)
    loss_D_tgt = 0.5 * soft_label_cross_entropy(model_D(tgt_fea.detach(), size), tgt_q1)
    return dict(loss_seg=loss_seg, loss_adv_tgt=loss_adv_tgt, loss_D_src=loss_D_src, loss_D_tgt=loss_D_tgt)


def fada_step(head: nn.Module, model_D: nn.Module, src_fea, tgt_fea, src_label,
              temperature: float = 1.8, ignore_index: int = 255):
    """One adversarial iteration after the backbone INCLUDING the backward passes, in the reference's order
    (aspp_fada.py:91-125; optimizer steps excluded, so all three D passes see the same D weights).
    Returns the four scalar losses."""
    size = src_label.shape[-2:]
    for p in list(head.parameters()) + list(model_D.parameters()):
        p.grad = None
    src_fea = src_fea.detach().requires_grad_(True)
    tgt_fea = tgt_fea.detach().requires_grad_(True)
    src_pred = head(src_fea, size).div(temperature)                                  # :92-94
    loss_seg = hard_cross_entropy(src_pred, src_label, ignore_index)
    loss_seg.backward()                                                              # :96
    src_soft = build_soft_label(src_pred.detach(), slot=0)                           # :99-100 (+ cat of :120)
    tgt_pred = head(tgt_fea, size).div(temperature)                                  # :103-104
    tgt_q0 = build_soft_label(tgt_pred.detach(), slot=0)                             # :105-108, cat of :111
    tgt_q1 = build_soft_label(tgt_pred.detach(), slot=1)                             # cat of :124
    loss_adv_tgt = 0.001 * soft_label_cross_entropy(model_D(tgt_fea, size), tgt_q0)  # :110-112
    loss_adv_tgt.backward()
    for p in model_D.parameters():                                                   # optimizer_D.zero_grad(), :117
        p.grad = None
    loss_D_src = 0.5 * soft_label_cross_entropy(model_D(src_fea.detach(), size), src_soft)   # :119-121
    loss_D_src.backward()
    loss_D_tgt = 0.5 * soft_label_cross_entropy(model_D(tgt_fea.detach(), size), tgt_q1)     # :123-125
    loss_D_tgt.backward()
    return loss_seg.detach(), loss_adv_tgt.detach(), loss_D_src.detach(), loss_D_tgt.detach()


def fada_iteration(head: nn.Module, model_D: nn.Module, optimizer_cls, optimizer_D, src_fea, tgt_fea, src_label,
                   temperature: float = 1.8, ignore_index: int = 255):
    """The same iteration WITH the optimizer steps where the reference takes them (aspp_fada.py:80-127 after the backbone;
    optimizer_fea belongs to the backbone and is out of scope): zero_grad :80-82, optimizer_cls.step() after the adversarial
    backward :114-115, optimizer_D.zero_grad() :117, optimizer_D.step() :127 -- so the two discriminator-update passes see the
    discriminator weights of the start of the iteration and the head's step does not feed back into this iteration's losses
    (the soft labels were taken before it).  Returns the four scalar losses."""
    size = src_label.shape[-2:]
    optimizer_cls.zero_grad()
    optimizer_D.zero_grad()
    src_fea = src_fea.detach().requires_grad_(True)
    tgt_fea = tgt_fea.detach().requires_grad_(True)
    src_pred = head(src_fea, size).div(temperature)
    loss_seg = hard_cross_entropy(src_pred, src_label, ignore_index)
    loss_seg.backward()
    src_soft = build_soft_label(src_pred.detach(), slot=0)
    tgt_pred = head(tgt_fea, size).div(temperature)
    tgt_q0 = build_soft_label(tgt_pred.detach(), slot=0)
    tgt_q1 = build_soft_label(tgt_pred.detach(), slot=1)
    loss_adv_tgt = 0.001 * soft_label_cross_entropy(model_D(tgt_fea, size), tgt_q0)
    loss_adv_tgt.backward()
    optimizer_cls.step()                                                             # :115
    optimizer_D.zero_grad()                                                          # :117
    loss_D_src = 0.5 * soft_label_cross_entropy(model_D(src_fea.detach(), size), src_soft)
    loss_D_src.backward()
    loss_D_tgt = 0.5 * soft_label_cross_entropy(model_D(tgt_fea.detach(), size), tgt_q1)
    loss_D_tgt.backward()
    optimizer_D.step()                                                               # :127
    return loss_seg.detach(), loss_adv_tgt.detach(), loss_D_src.detach(), loss_D_tgt.detach()
