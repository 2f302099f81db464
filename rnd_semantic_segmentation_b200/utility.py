"""Drop-in replacements for the hot-path helpers of core/utils/utility.py.

Same names, argument meaning and return conventions as the reference:
``soft_label_cross_entropy`` (:172-177), ``inference`` (:179-191), ``intersectionAndUnionGPU``
(:148-161), ``intersectionAndUnion`` (:133-145), ``confusion_matrix`` (:347-359), ``AverageMeter``
(:24-72) -- plus the fused entry point ``segmentation_eval_step`` that the reference's tester loop
(core/testers/aspp_tester.py:57-74) collapses into.
"""
from __future__ import annotations

import weakref
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib, ops
from .ops import soft_label_cross_entropy  # noqa: F401  (re-exported under the reference's name)

__all__ = ["soft_label_cross_entropy", "inference", "multi_scale_inference", "intersectionAndUnion", "intersectionAndUnionGPU",
           "confusion_matrix", "AverageMeter", "segmentation_eval_step", "iutr_from_confusion", "LazyProbabilities", "BatchedEvaluator",
           "pseudo_label_map", "get_color_palette", "save_pseudo_label"]


# ------------------------------------------------------------------------------------------------
# AverageMeter (utility.py:24-72) -- host-side scalar bookkeeping, numerically identical
# ------------------------------------------------------------------------------------------------
class AverageMeter(object):
    """Accumulates per-frame intersection / union / target / output areas and IoU / F1 sums."""

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
        iou = intersection / (union + 1e-10)
        f1 = 2 * intersection / (target + res + 1e-10)
        self.intersection_sum += intersection
        self.union_sum += union
        self.target_sum += target
        self.res_sum += res
        self.count += 1
        self.iou_sum += iou
        self.f1_sum += f1

    def results(self):
        macro_f1 = self.f1_sum / float(self.count)
        macro_iou = self.iou_sum / float(self.count)
        micro_f1 = 2 * self.intersection_sum / (self.target_sum + self.res_sum + 1e-10)
        micro_iou = self.intersection_sum / (self.union_sum + 1e-10)
        return dict(macro_iou=macro_iou, macro_f1=macro_f1, micro_iou=micro_iou, micro_f1=micro_f1)

    def summary(self, logger, num_classes=2):
        r = self.results()
        logger.info('Macro metric, val result: mIoU/mF1 {:.4f}/{:.4f}.'.format(np.mean(r["macro_iou"]), np.mean(r["macro_f1"])))
        logger.info('Micro metric, val result: mIoU/mF1 {:.4f}/{:.4f}.'.format(np.mean(r["micro_iou"]), np.mean(r["micro_f1"])))
        for i in range(num_classes):
            logger.info('Macro metric, class {} iou/f1 score: {:.4f}/{:.4f}.'.format(i, r["macro_iou"][i], r["macro_f1"][i]))
            logger.info('Micro metric, class {} iou/f1 score: {:.4f}/{:.4f}.'.format(i, r["micro_iou"][i], r["micro_f1"][i]))


# ------------------------------------------------------------------------------------------------
# fused eval step
# ------------------------------------------------------------------------------------------------
def iutr_from_confusion(cm: torch.Tensor):
    """(intersection, union, target, output) areas of utility.py:148-161 from a confusion matrix whose rows
    exclude ignored truth: I = diag, T = row sums, O = column sums, U = O + T - I (exact int64)."""
    inter = torch.diagonal(cm, dim1=-2, dim2=-1)
    tgt = cm.sum(dim=-1)
    out = cm.sum(dim=-2)
    return inter, out + tgt - inter, tgt, out


def segmentation_eval_step(logits_lr: torch.Tensor, labels: torch.Tensor, ignore_index: int = 255,
                           cm: Optional[torch.Tensor] = None, per_frame: bool = False, want_pred=False):
    """One launch for the whole post-head part of aspp_tester.py:60-72: align-corners upsample to
    ``labels.shape[-2:]``, softmax, argmax (first index on ties, bit-exact with the reference on CUDA)
    and the C x C int64 confusion matrix accumulated into ``cm``.  Returns (cm, pred or None)."""
    labels = _lib.as_label_tensor(labels)              # int64 or uint8 consumed as they are
    return _lib.upsample_argmax_confusion(logits_lr.float().contiguous(), labels.contiguous(), labels.shape[-2:],
                                          ignore_index=ignore_index, cm=cm, per_frame=per_frame, want_pred=want_pred)


class BatchedEvaluator:
    """The tester loop of aspp_tester.py:57-72 after the backbone, with the upsample + argmax + confusion-matrix kernel (K4) run
    over SEVERAL frames per launch: ``step(features, labels)`` runs the head for the frame (TEST.BATCH_SIZE = 1, as the reference)
    and queues its low-res logits (2.5 MB) with the frame's labels (int64 or uint8, used in place -- no copy); every ``frames``
    steps one launch processes the queued frames.  A single 1024 x 2048 frame is too little work to fill 148 SMs with tall tiles
    (34 us per frame launched alone, 19 us per frame at 8 per launch), so batching the launches is what the eval img/s needs.
    ``finish()`` flushes the queue and returns the int64 [C,C] matrix (with ``per_frame=True`` also the list of per-frame
    matrices, for the macro metrics of AverageMeter).  Bit-identical to one ``segmentation_eval_step`` per frame."""

    def __init__(self, classifier, num_classes: int, ignore_index: int = 255, frames: int = 8, per_frame: bool = False, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.classifier, self.ignore_index, self.frames, self.per_frame = classifier, ignore_index, max(1, int(frames)), per_frame
        self.cm = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=dev)
        self.frame_cms = []
        self._logits, self._labels = [], []

    def step(self, features: torch.Tensor, labels: torch.Tensor):
        with torch.no_grad():
            lg = self.classifier.logits(features)
        return self.step_logits(lg, labels)

    def step_logits(self, logits_lr: torch.Tensor, labels: torch.Tensor):
        n = logits_lr.shape[0]
        labels = _lib.as_label_tensor(labels).reshape(n, *labels.shape[-2:])
        for f in range(n):
            self._logits.append(logits_lr[f].detach().float().contiguous())
            self._labels.append(labels[f].contiguous())
        if len(self._logits) >= self.frames:
            self.flush()
        return self.cm

    def flush(self):
        if not self._logits:
            return
        if self.per_frame:
            cms, _ = _lib.upsample_argmax_confusion_frames(self._logits, self._labels, self._labels[0].shape[-2:],
                                                           ignore_index=self.ignore_index, per_frame=True)
            self.frame_cms.extend(cms.unbind(0))
            self.cm += cms.sum(0)
        else:
            _lib.upsample_argmax_confusion_frames(self._logits, self._labels, self._labels[0].shape[-2:],
                                                  ignore_index=self.ignore_index, cm=self.cm)
        self._logits, self._labels = [], []

    def finish(self):
        self.flush()
        return (self.cm, self.frame_cms) if self.per_frame else self.cm


# A prediction produced by LazyProbabilities.max remembers the confusion matrix computed in the same
# launch, so the reference tester's follow-up calls confusion_matrix(...) / intersectionAndUnionGPU(...)
# on that very tensor are answered without touching the pixels again.
_pred_cache = weakref.WeakValueDictionary()


class _PredRecord:
    """What LazyProbabilities.max() remembers about the prediction it returned: identity (storage address, element count),
    the version counters of the prediction and of the labels at that moment, and the matrix of the same launch.  The record
    holds NO reference to the prediction tensor (the tensor holds the record, not the other way round: no reference cycle,
    the frame's tensors are freed by refcount as soon as the tester drops them)."""
    __slots__ = ("pred_numel", "pred_version", "labels_ptr", "labels_version", "labels_ref", "cm", "ignore_index", "__weakref__")


class LazyProbabilities:
    """What ``inference`` / ``multi_scale_inference`` return: stands for the [1,C,H,W] probability tensor of the reference
    (softmax of the upsampled logits; with test-time augmentation the average over the flipped / rescaled members) without
    materialising it.  ``.max(1)`` gives (None, int64 argmax) through the fused kernels -- K4 for a single member, K7
    (``tta_argmax_confusion``) for an ensemble -- together with the confusion matrix of the same launch; any other use
    materialises the real tensor (``.materialize()`` / ``torch.as_tensor`` semantics).

    ``members``: low-res logits [1,C,h_m,w_m]; ``flips[m]``: member m saw the mirrored image; ``divisors``: the scalar
    divisions the reference applies to the sum, in order (utility.py:188, :207-209)."""

    def __init__(self, logits_lr, label, ignore_index=255, members=None, flips=None, divisors=()):
        self.members = [logits_lr] if members is None else list(members)
        self.flips = [False] * len(self.members) if flips is None else [bool(f) for f in flips]
        self.divisors = tuple(float(d) for d in divisors)
        self.logits_lr = self.members[0]
        self.label = label
        self.ignore_index = ignore_index
        self._full = None

    @property
    def is_ensemble(self):
        return len(self.members) > 1 or any(self.flips) or len(self.divisors) > 0

    @property
    def shape(self):
        n, c = self.logits_lr.shape[:2]
        return torch.Size((n, c) + tuple(self.label.shape[-2:]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def _members32(self):
        return [m.detach().float().contiguous() for m in self.members]

    def materialize(self) -> torch.Tensor:
        """The real [1,C,H,W] probability tensor, written by the fused kernel (K7 with ``want_probs``: upsample + ATen's exact
        soft-max sequence per label pixel, bit-equal to ``F.softmax(F.interpolate(...))`` on CUDA) -- single member or ensemble."""
        if self._full is None:
            _, _, probs = _lib.tta_argmax_confusion(self._members32(), self.flips, self.label.shape[-2:],
                                                    divisors=self.divisors, want_probs=True)
            self._full = probs.unsqueeze(0)
        return self._full

    def max(self, dim=None, keepdim=False):
        if dim != 1 or keepdim:
            return self.materialize().max(dim, keepdim) if dim is not None else self.materialize().max()
        labels = _lib.as_label_tensor(self.label)
        labels = labels.reshape(self.logits_lr.shape[0], *labels.shape[-2:]).contiguous()
        if self.is_ensemble:
            cm, pred, _ = _lib.tta_argmax_confusion(self._members32(), self.flips, labels.shape[-2:], labels=labels,
                                                    divisors=self.divisors, ignore_index=self.ignore_index, want_pred=True)
            pred = pred.unsqueeze(0)
        else:
            cm, pred = _lib.upsample_argmax_confusion(self.logits_lr, labels, labels.shape[-2:],
                                                      ignore_index=self.ignore_index, per_frame=False, want_pred=True)
        rec = _PredRecord()
        rec.pred_numel, rec.pred_version = pred.numel(), pred._version
        # (the reshaped ``labels`` is a temporary view; views share storage and version counter with the caller's tensor)
        rec.labels_ptr, rec.labels_version, rec.labels_ref = labels.data_ptr(), labels._version, weakref.ref(self.label)
        rec.cm, rec.ignore_index = cm, self.ignore_index
        _pred_cache[pred.data_ptr()] = rec
        pred._b200seg_record = rec          # keeps the record alive exactly as long as the prediction tensor
        return None, pred

    def argmax(self, dim=None, keepdim=False):
        if dim == 1 and not keepdim:
            return self.max(1)[1]
        return self.materialize().argmax(dim, keepdim)

    def label_map(self, dtype=torch.uint8) -> torch.Tensor:
        """[H,W] argmax labels written by the fused kernel directly in ``dtype`` (uint8: the pseudo-label format of
        aspp_tester.py:40-45 -- 1 byte per pixel leaves the kernel instead of an int64 map that is narrowed afterwards).  No
        confusion matrix, the ground truth is not read."""
        size = self.label.shape[-2:]
        if self.is_ensemble:
            _, pred, _ = _lib.tta_argmax_confusion(self._members32(), self.flips, size, divisors=self.divisors, want_pred=dtype)
            return pred
        _, pred = _lib.upsample_argmax_confusion(self.logits_lr.detach().float().contiguous(), None, size, want_pred=dtype)
        return pred[0]

    def cpu(self):
        return self.materialize().cpu()

    def __getitem__(self, idx):
        return self.materialize()[idx]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):   # any torch.* call on the lazy object
        kwargs = kwargs or {}
        args = tuple(a.materialize() if isinstance(a, LazyProbabilities) else a for a in args)
        return func(*args, **kwargs)


def inference(feature_extractor, classifier, image, label, flip=True):
    """utility.py:179-191.  Returns a LazyProbabilities standing for the reference's [1,C,H,W] tensor: flip=False (what
    ASPPTester uses, aspp_tester.py:60) is one member; flip=True is the two-member ensemble
    (softmax(up(head(image))) + mirror(softmax(up(head(mirror(image)))))) / 2, evaluated per label pixel by K7."""
    if flip:
        image = torch.cat([image, torch.flip(image, [3])], 0)
    with torch.no_grad():
        output = classifier(feature_extractor(image))
    if not flip:
        return LazyProbabilities(output[:1].contiguous(), label)
    return LazyProbabilities(None, label, members=[output[0:1].contiguous(), output[1:2].contiguous()], flips=[False, True],
                             divisors=(2,))


def multi_scale_inference(feature_extractor, classifier, image, label, flip=True, scales=[0.7, 1.0, 1.3]):
    """utility.py:193-209: the backbone + head run once per scale (and once more on the mirrored image when ``flip``); the
    reference then upsamples, soft-maxes, un-mirrors and sums 2 * len(scales) full-resolution tensors and divides by
    len(scales) and by 2.  Here the low-res logits of all members go to ONE K7 launch, summed in the reference's order."""
    size = image.shape[-2:]
    members, flips = [], []
    with torch.no_grad():
        for s in scales:
            x = F.interpolate(image, size=(int(size[0] * s), int(size[1] * s)), mode='bilinear', align_corners=True)
            members.append(classifier(feature_extractor(x))[:1].contiguous())
            flips.append(False)
            if flip:
                members.append(classifier(feature_extractor(torch.flip(x, [3])))[:1].contiguous())
                flips.append(True)
    if len(members) > 8:
        raise _lib.B200SegError(f"multi_scale_inference: {len(members)} ensemble members, at most 8 are supported")
    return LazyProbabilities(None, label, members=members, flips=flips, divisors=(len(scales), 2) if flip else (len(scales),))


def pseudo_label_map(output) -> np.ndarray:
    """H x W uint8 label map of a (lazy or real) [1,C,H,W] probability tensor: what ``save_distill`` (aspp_tester.py:40-42)
    derives with ``output.cpu().numpy().squeeze().argmax(0)`` -- but through the fused argmax, so 2 MB of labels cross PCIe
    instead of the 159 MB probability tensor (first index on ties in both)."""
    if isinstance(output, LazyProbabilities):
        return output.label_map(torch.uint8).cpu().numpy()
    pred = torch.as_tensor(output).max(1)[1]
    return pred.reshape(pred.shape[-2:]).to(torch.uint8).cpu().numpy()


def get_color_palette(pred, palette):
    """utility.py:211-217: ``pred`` H x W numpy array with label ids -> PIL image in mode 'P' carrying ``palette``."""
    from PIL import Image
    label = Image.fromarray(np.asarray(pred).astype('uint8')).convert('P')
    label.putpalette(palette)
    return label


def save_pseudo_label(output, path: str, palette):
    """The body of ``ASPPTester.save_distill`` (aspp_tester.py:33-45) after the folder handling: argmax -> palette PNG."""
    mask = get_color_palette(pseudo_label_map(output), palette)
    mask.save(path)
    return mask


def _cached_record(pd: torch.Tensor, gt: torch.Tensor):
    """The record of ``pd`` if it is still the untouched prediction of LazyProbabilities.max() against these very labels:
    same storage, same element count, and NEITHER tensor edited in place since (version counters; views share them, so
    ``pred.flatten()`` / ``y.flatten()`` as the tester passes them qualify).  The reference recounts from the tensors it is
    given, so any in-place edit (id remapping, masking) must invalidate the memo.  intersectionAndUnionGPU's own raw-pointer
    write of ignore_index does not bump the version and only touches pixels the matrix skips anyway."""
    rec = _pred_cache.get(pd.data_ptr())
    if rec is None or rec.pred_numel != pd.numel() or rec.pred_version != pd._version:
        return None
    if gt.data_ptr() != rec.labels_ptr or gt.numel() != pd.numel():
        return None
    lab = rec.labels_ref()
    if lab is None or lab._version != rec.labels_version or gt._version != rec.labels_version:
        return None
    return rec


class _Cfg:
    pass


def confusion_matrix(cfg, pd, gt):
    """utility.py:347-359: int64 [C,C], rows = truth, cols = prediction, truth == 255 skipped; returned on the
    CPU like the reference (the tester adds it to a CPU accumulator, aspp_tester.py:69)."""
    num_classes = cfg.MODEL.NUM_CLASSES if not isinstance(cfg, int) else cfg
    if not pd.is_cuda:
        raise _lib.B200SegError("confusion_matrix: expected CUDA tensors (b200seg has no CPU fallback)")
    rec = _cached_record(pd, gt)
    if rec is not None and rec.ignore_index == 255 and rec.cm.shape[-1] == num_classes:
        return rec.cm.cpu()
    cm = _lib.confusion_from_pred(_lib.as_label_tensor(pd.reshape(-1)).contiguous(), _lib.as_label_tensor(gt.reshape(-1)).contiguous(),
                                  num_classes, 255)
    return cm.cpu()


def intersectionAndUnionGPU(output, target, K, ignore_index=255):
    """utility.py:148-161: float32 [K] x4 on the GPU; writes ignore_index into ``output`` where the target is
    ignored (the reference mutates its argument through a view, :154)."""
    assert output.dim() in [1, 2, 3]
    assert output.shape == target.shape
    if not output.is_cuda:
        raise _lib.B200SegError("intersectionAndUnionGPU: expected CUDA tensors (b200seg has no CPU fallback)")
    out_flat = output.view(-1)
    tgt_flat = _lib.as_label_tensor(target.reshape(-1)).contiguous()
    if out_flat.dtype not in _lib.LABEL_DTYPES:
        work = out_flat.long().contiguous()
        cm = _lib.confusion_from_pred(work, tgt_flat, K, ignore_index, mutate_pd=True)
        out_flat.copy_(work.to(out_flat.dtype))
    else:
        cm = _lib.confusion_from_pred(out_flat, tgt_flat, K, ignore_index, mutate_pd=True)       # int64 / uint8 maps, any mix
    i, u, t, o = iutr_from_confusion(cm)
    return i.float(), u.float(), t.float(), o.float()


def intersectionAndUnion(output, target, K, ignore_index=255):
    """numpy twin, utility.py:133-145 (host-side helper kept for API completeness): areas of
    intersection / union / target / output per class, ignored-target pixels excluded."""
    assert output.ndim in [1, 2, 3]
    assert output.shape == target.shape
    pd = np.asarray(output).reshape(-1).astype(np.int64)
    gt = np.asarray(target).reshape(-1).astype(np.int64)
    keep = gt != ignore_index

    def count(v):
        v = v[(v >= 0) & (v < K)]
        return np.bincount(v, minlength=K)[:K]

    area_output = count(pd[keep])
    area_target = count(gt)
    area_intersection = count(pd[keep & (pd == gt)])
    return area_intersection, area_output + area_target - area_intersection, area_target, area_output
