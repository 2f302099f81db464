"""numpy restatement of the ATen algorithms underneath the reference's hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Independent of torch, so it
cross-checks ``torch_oracle`` and documents the exact arithmetic our kernels
implement.  float32 throughout unless stated; formulas follow
``ATen/native/UpSample.h`` (area_pixel_compute_scale / _source_index) and
``ATen/native/cuda/UpSample.cuh`` + ``UpSampleBilinear2d.cu`` for the bilinear
expression form, ``SoftMax.cu`` (spatial softmax: max, sum of exp(x-max) in
class order, exp(x-max)/sum) and ``Loss.cpp``/``NLLLoss2d`` for the loss.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def ac_scale(in_size: int, out_size: int) -> np.float32:
    """align_corners=True scale: (in-1)/(out-1) in float32, 0 when out == 1."""
    if out_size > 1:
        return F32(F32(in_size - 1) / F32(out_size - 1))
    return F32(0)


def ac_taps(in_size: int, out_size: int):
    """Per output index: (i0, i1, lambda0, lambda1) exactly as ATen computes them.

    src = scale * dst (float32); i0 = (int)src; i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1.
    """
    scale = ac_scale(in_size, out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = (scale * dst).astype(np.float32)
    i0 = src.astype(np.int32)
    i1 = i0 + (i0 < in_size - 1).astype(np.int32)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (F32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def upsample_bilinear_ac(x: np.ndarray, size) -> np.ndarray:
    """x [N,C,h,w] float32 -> [N,C,H,W]; expression form
    l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11) in float32 (no FMA contraction
    here -- numpy rounds every product; the GPU kernels may differ by an ulp)."""
    x = np.asarray(x, dtype=np.float32)
    H, W = size
    h, w = x.shape[-2:]
    y0, y1, ly0, ly1 = ac_taps(h, H)
    x0, x1, lx0, lx1 = ac_taps(w, W)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    t = (lx0 * top[..., x0]).astype(F32) + (lx1 * top[..., x1]).astype(F32)
    b = (lx0 * bot[..., x0]).astype(F32) + (lx1 * bot[..., x1]).astype(F32)
    ly0 = ly0[:, None]
    ly1 = ly1[:, None]
    return ((ly0 * t).astype(F32) + (ly1 * b).astype(F32)).astype(F32)


def log_softmax(x: np.ndarray, axis: int = 1) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=axis, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=axis, keepdims=True))


def hard_cross_entropy(logits: np.ndarray, labels: np.ndarray, ignore_index: int = 255):
    """float64 evaluation of mean_{valid}( -log_softmax(x)[y] ); returns (loss, n_valid)."""
    ls = log_softmax(logits, 1)
    valid = labels != ignore_index
    safe = np.where(valid, labels, 0)
    picked = np.take_along_axis(ls, safe[:, None], axis=1)[:, 0]
    n_valid = int(valid.sum())
    total = -(picked * valid).sum()
    return (total / n_valid if n_valid else float("nan")), n_valid


def hard_cross_entropy_grad(logits: np.ndarray, labels: np.ndarray, ignore_index: int = 255) -> np.ndarray:
    """d loss / d logits = (softmax - onehot)/n_valid on valid pixels, 0 elsewhere (float64)."""
    p = np.exp(log_softmax(logits, 1))
    valid = labels != ignore_index
    safe = np.where(valid, labels, 0)
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, safe[:, None], 1.0, axis=1)
    g = (p - onehot) * valid[:, None]
    return g / max(int(valid.sum()), 1)


def upsample_bilinear_ac_adjoint(g: np.ndarray, in_hw) -> np.ndarray:
    """Adjoint of upsample_bilinear_ac: g [N,C,H,W] -> [N,C,h,w] (float64 accumulation)."""
    g = np.asarray(g, dtype=np.float64)
    h, w = in_hw
    H, W = g.shape[-2:]
    y0, y1, ly0, ly1 = ac_taps(h, H)
    x0, x1, lx0, lx1 = ac_taps(w, W)
    out = np.zeros(g.shape[:-2] + (h, w), dtype=np.float64)
    rows = np.zeros(g.shape[:-2] + (h, W), dtype=np.float64)
    np.add.at(rows, (Ellipsis, y0, slice(None)), g * ly0[:, None].astype(np.float64))
    np.add.at(rows, (Ellipsis, y1, slice(None)), g * ly1[:, None].astype(np.float64))
    np.add.at(out, (Ellipsis, x0), rows * lx0.astype(np.float64))
    np.add.at(out, (Ellipsis, x1), rows * lx1.astype(np.float64))
    return out


def soft_label_cross_entropy(pred: np.ndarray, soft: np.ndarray, weights=None) -> float:
    per_px = -(np.asarray(soft, np.float64) * log_softmax(pred, 1)).sum(axis=1)
    if weights is not None:
        per_px = per_px * weights
    return float(per_px.mean())


def soft_label_cross_entropy_grad(pred: np.ndarray, soft: np.ndarray, weights=None) -> np.ndarray:
    """(softmax(p) * sum_c q_c - q) * w / (N*H*W)."""
    q = np.asarray(soft, np.float64)
    p = np.exp(log_softmax(pred, 1))
    g = p * q.sum(axis=1, keepdims=True) - q
    if weights is not None:
        g = g * np.asarray(weights, np.float64)[:, None]
    n, _, hh, ww = pred.shape
    return g / (n * hh * ww)


def softmax_first_max(x: np.ndarray) -> np.ndarray:
    """argmax over axis 1 of the float32 softmax as ATen's spatial softmax computes it:
    m = max_c x; s = sum_c expf(x_c - m) accumulated in class order in float32;
    p_c = expf(x_c - m) / s; first index attaining max_c p_c.
    (np.exp on float32 is not bit-identical to CUDA expf; exactness claims are made
    against torch CUDA on the GPU box, this is the documented algorithm.)"""
    x = np.asarray(x, dtype=np.float32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp((x - m).astype(F32)).astype(F32)
    s = np.zeros(e.shape[:1] + e.shape[2:], dtype=F32)
    for c in range(e.shape[1]):
        s = (s + e[:, c]).astype(F32)
    p = (e / s[:, None]).astype(F32)
    return p.argmax(axis=1).astype(np.int64)


def confusion_matrix(num_classes: int, pd: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """cmt[gt, pd] += 1 for gt != 255 (utility.py:347-359), int64 [C,C]."""
    pd = np.asarray(pd).reshape(-1).astype(np.int64)
    gt = np.asarray(gt).reshape(-1).astype(np.int64)
    keep = gt != 255
    return np.bincount(gt[keep] * num_classes + pd[keep],
                       minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def iutr_from_confusion(cm: np.ndarray):
    """intersection / union / target / output areas (utility.py:133-145) derived from one
    frame's confusion matrix whose rows exclude ignored truth: I = diag, T = row sums,
    O = column sums, U = O + T - I."""
    cm = np.asarray(cm, dtype=np.int64)
    inter = np.diag(cm).copy()
    tgt = cm.sum(axis=1)
    out = cm.sum(axis=0)
    return inter, out + tgt - inter, tgt, out


def aspp_head(x: np.ndarray, weights, biases, rates=(6, 12, 18, 24)) -> np.ndarray:
    """Direct (slow, float64) evaluation of classifier.py:26-29 for tiny shapes:
    out[n,c,y,x] = sum_r ( b_r[c] + sum_{ci,ky,kx} W_r[c,ci,ky,kx] * X[n,ci,y+(ky-1)r,x+(kx-1)r] )."""
    x = np.asarray(x, dtype=np.float64)
    n, cin, h, w = x.shape
    cout = weights[0].shape[0]
    out = np.zeros((n, cout, h, w), dtype=np.float64)
    for wgt, b, r in zip(weights, biases, rates):
        wgt = np.asarray(wgt, np.float64)
        out += np.asarray(b, np.float64)[None, :, None, None]
        for ky in range(3):
            for kx in range(3):
                dy, dx = (ky - 1) * r, (kx - 1) * r
                ys0, ys1 = max(0, -dy), min(h, h - dy)
                xs0, xs1 = max(0, -dx), min(w, w - dx)
                if ys0 >= ys1 or xs0 >= xs1:
                    continue
                patch = x[:, :, ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
                out[:, :, ys0:ys1, xs0:xs1] += np.einsum("oc,nchw->nohw", wgt[:, :, ky, kx], patch)
    return out


# --------------------------------------------------------------------------
# 8f-3  test-time augmentation from the members' low-res logits
#        (core/utils/utility.py:179-191 flip=True, :193-209 multi_scale_inference)
# --------------------------------------------------------------------------
def softmax_probabilities(x: np.ndarray) -> np.ndarray:
    """float32 spatial softmax over axis 1, ATen's sequence (see softmax_first_max)."""
    x = np.asarray(x, dtype=np.float32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp((x - m).astype(F32)).astype(F32)
    s = np.zeros(e.shape[:1] + e.shape[2:], dtype=F32)
    for c in range(e.shape[1]):
        s = (s + e[:, c]).astype(F32)
    return (e / s[:, None]).astype(F32)


def tta_probabilities(members, flips, size, divisors=()) -> np.ndarray:
    """sum over members, in order, of the (un-mirrored) softmax of the upsampled logits [1,C,h_m,w_m]; then the scalar
    divisions in order (float32 IEEE divisions, as ATen's CPU kernel performs ``tensor / python_scalar``)."""
    total = None
    for lg, fl in zip(members, flips):
        pr = softmax_probabilities(upsample_bilinear_ac(np.asarray(lg, dtype=F32), size))
        if fl:
            pr = pr[..., ::-1]
        total = pr.copy() if total is None else (total + pr).astype(F32)
    for d in divisors:
        total = (total / F32(d)).astype(F32)
    return total


# --------------------------------------------------------------------------
# 8f-4  optimizer steps, the arithmetic of torch/optim/sgd.py::_single_tensor_sgd and torch/optim/adam.py::_single_tensor_adam
#        as the reference configures them (aspp_trainer.py:25-26, fada_adapter.py:24).  ``a + alpha * b`` ops are single
#        FMAs in ATen; numpy has no fma, so they are evaluated in float64 and rounded once to float32 (the float64 product of
#        two float32 values is exact; the rare double rounding of the sum is below the 1e-6 parity bar).
# --------------------------------------------------------------------------
def _fma32(a, b, c):
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64) + np.asarray(c, dtype=np.float64)).astype(F32)


def sgd_step(p, g, buf, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, grad_scale=1.0):
    """One torch.optim.SGD update.  ``buf`` is None on the first step (torch's momentum_buffer is None).  Returns (p, buf)."""
    p, g = np.asarray(p, dtype=F32), np.asarray(g, dtype=F32)
    if grad_scale != 1.0:
        g = (g * F32(grad_scale)).astype(F32)
    if weight_decay != 0:
        g = _fma32(F32(weight_decay), p, g)                       # grad.add(param, alpha=weight_decay)
    d = g
    if momentum != 0:
        if buf is None:
            buf = g.copy()                                       # torch.clone(grad)
        else:
            buf = _fma32(F32(1.0 - dampening), g, (np.asarray(buf, dtype=F32) * F32(momentum)).astype(F32))   # mul_().add_(grad, alpha)
        d = _fma32(F32(momentum), buf, g) if nesterov else buf
    return _fma32(F32(-lr), d, p), buf                           # param.add_(grad, alpha=-lr)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """One torch.optim.Adam update (no amsgrad); ``step`` counts this update (>= 1); m, v start at zero.  Returns (p, m, v)."""
    p, g, m, v = (np.asarray(t, dtype=F32) for t in (p, g, m, v))
    if grad_scale != 1.0:
        g = (g * F32(grad_scale)).astype(F32)
    if weight_decay != 0:
        g = _fma32(F32(weight_decay), p, g)
    m = _fma32(F32(1.0 - beta1), (g - m).astype(F32), m)          # exp_avg.lerp_(grad, 1 - beta1)
    v = _fma32((F32(1.0 - beta2) * g).astype(F32), g, (v * F32(beta2)).astype(F32))   # mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2_sqrt = (1.0 - beta2 ** step) ** 0.5
    denom = ((np.sqrt(v).astype(F32) / F32(bc2_sqrt)).astype(F32) + F32(eps)).astype(F32)
    p = _fma32(F32(-(lr / bc1)), (m / denom).astype(F32), p)      # addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v
