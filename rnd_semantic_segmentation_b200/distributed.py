"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on the
B200 box, gloo in the CPU tests).  The hot path shards by image / frame with no data-path collective;
the only exchange steps are

* training: MEAN all-reduce of the head / discriminator gradients each step (DDP semantics of
  train_distill.py:54-62; train_adv.py shards data but never syncs) -- one flat
  bucket per module, issued on the stream that produced the gradients;
* eval: SUM all-reduce of the int64 C x C confusion matrix (order independent => bit-exact at any world size).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: Optional[str] = None) -> bool:
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* when launched by torchrun."""
    rank, world, local = env_rank_world()
    if world <= 1:
        return False
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, init_method="env://")
    return True


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Frame i -> rank i mod world."""
    return list(range(rank, n_items, world))


_symm_warned = [False]


def symmetric_flat_buffer(numel: int, device, group=None):
    """(buffer, handle) -- ``numel`` fp32 (rounded up to a multiple of 4) in NVLink symmetric memory, mapped by every rank of
    ``group`` (torch's symmetric-memory rendezvous: plumbing for the peer / multicast / signal-pad addresses our own all-reduce
    kernel works on, csrc/p2p_allreduce.cu) -- or None when that is not available (CPU / gloo, single rank, B200SEG_P2P_ALLREDUCE=0,
    no peer access).  A collective call: every rank must make it, in the same order."""
    if os.environ.get("B200SEG_P2P_ALLREDUCE", "1") == "0" or not is_distributed():
        return None
    if torch.device(device).type != "cuda" or dist.get_backend(group) != "nccl":
        return None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        n4 = (int(numel) + 3) // 4 * 4
        buf = symm_mem.empty(n4, dtype=torch.float32, device=torch.device(device))
        hdl = symm_mem.rendezvous(buf, group=group if group is not None else dist.group.WORLD)
        buf.zero_()
        return buf, hdl
    except Exception as e:                                  # pragma: no cover  (depends on the box)
        if not _symm_warned[0]:
            _symm_warned[0] = True
            print(f"b200seg: symmetric memory unavailable ({e!r}); gradient buckets fall back to NCCL", flush=True)
        return None


class FlatGradBucket:
    """One flat fp32 buffer holding all gradients of a module; parameters' .grad are views into it, so the
    step's all-reduce is ONE collective (head: 1 400 908 fp32 = 5.6 MB; discriminator: 20.2 MB).  ``symmetric``: place the buffer
    in NVLink symmetric memory so that the all-reduce can be our own peer-memory kernel instead of NCCL."""

    def __init__(self, params: Iterable[torch.nn.Parameter], symmetric: bool = False, group=None):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self._symm = symmetric_flat_buffer(total, dev, group) if symmetric else None
        if self._symm is not None:
            self._flat_full = self._symm[0]                  # padded to 4 floats; the tail stays zero
            self.flat = self._flat_full[:total]
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        self.views = []
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()

    def gather_grads(self):
        """Bring the parameters' current .grad into the flat buffer (one concatenation kernel) and alias them to it."""
        if all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self.params, self.views)):
            return
        pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        torch.cat(pieces, out=self.flat)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def p2p_allreduce_mean_(self, blocks: int = 8):
        """Our own mean all-reduce over peer memory (csrc/p2p_allreduce.cu), enqueued on the CURRENT stream; every rank must call it."""
        from . import _lib
        buf, hdl = self._symm
        _lib.p2p_allreduce_mean(hdl.buffer_ptrs, getattr(hdl, "multicast_ptr", 0) or 0, hdl.signal_pad_ptrs, hdl.rank, hdl.world_size,
                                buf.numel(), blocks=blocks, device=buf.device)

    def allreduce_mean_(self, group=None):
        self.gather_grads()
        if self._symm is not None:
            self.p2p_allreduce_mean_()
            return self.flat
        if is_distributed():
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)      # mean inside the collective
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        return self.flat


class OverlappedGradBucket(FlatGradBucket):
    """Flat gradient bucket whose MEAN all-reduce (and, optionally, the optimizer step behind it) runs on the bucket's own
    stream while the compute stream carries on -- DDP semantics (train_distill.py:54-62) for a module the reference never
    synchronises (train_adv.py shards the data but wraps nothing).

    Use for the PixelDiscriminator (5 057 702 fp32 = 20.2 MB): ``zero_grads_()`` instead of ``optimizer_D.zero_grad()``
    (aspp_fada.py:117) keeps the parameters' .grad aliased to the bucket, so the two discriminator backward passes (:119-125)
    accumulate straight into it; ``begin_allreduce_()`` after the second one starts the collective; ``step(optimizer_D)``
    enqueues the Adam update behind it on the same stream; ``wait()`` -- before the discriminator is used again, i.e. after the
    NEXT iteration's head forward / backward -- joins.  The collective and the update therefore overlap the next iteration's
    source-domain head step instead of sitting between iterations."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, p2p_blocks: int = 8):
        super().__init__(params, symmetric=True, group=group)
        self.group = group
        self.p2p_blocks = int(p2p_blocks)
        self._pending = False
        self.stream = torch.cuda.Stream(device=self.flat.device) if self.flat.is_cuda else None
        self.ready_event = torch.cuda.Event() if self.flat.is_cuda else None
        if is_distributed() and self.flat.is_cuda:           # first launches / communicator channels now, not in the first steps
            with torch.cuda.stream(self.stream):
                for _ in range(2):
                    if self._symm is not None:
                        self.p2p_allreduce_mean_(self.p2p_blocks)
                    else:
                        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            torch.cuda.current_stream().wait_stream(self.stream)
            self.flat.zero_()

    def zero_grads_(self):
        """optimizer.zero_grad() that keeps .grad aliased to the flat buffer (so backward accumulates in place)."""
        self.wait()
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def begin_allreduce_(self):
        self.gather_grads()                                   # no-op when .grad already aliases the bucket
        if not is_distributed():
            return
        if self.stream is None:                               # CPU tensors (gloo tests): blocking
            self.allreduce_mean_(self.group)
            return
        self.ready_event.record()
        self.stream.wait_event(self.ready_event)
        with torch.cuda.stream(self.stream):
            if self._symm is not None:
                self.p2p_allreduce_mean_(self.p2p_blocks)
            elif dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(dist.get_world_size(self.group))
        self._pending = True

    def step(self, optimizer):
        if not self._pending or self.stream is None:
            return optimizer.step()
        with torch.cuda.stream(self.stream):
            return optimizer.step()

    def wait(self):
        if self._pending and self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._pending = False


class HeadGradBucket(FlatGradBucket):
    """Flat gradient bucket of an ASPP_Classifier_V2 whose all-reduce OVERLAPS the head's data-gradient GEMM.

    ``head.forward_loss(x, labels, grad_bucket=bucket)`` makes the backward write the weight gradients straight into the
    bucket (no concatenation pass), record ``ready_event`` right after the weight-gradient GEMM, and call
    ``begin_allreduce_()``, which runs the MEAN all-reduce on the bucket's own stream while the compute stream goes on
    with the data-gradient GEMM.  ``wait()`` joins the two streams before the optimizer (or anything else) reads .grad.
    Gradients are overwritten each step, not accumulated (zero_grad-every-step semantics of the reference trainers)."""

    def __init__(self, head: torch.nn.Module, group=None, overlap_ctas: Optional[int] = None):
        # CTAs of the overlapped all-reduce = SMs the dgrad GEMM leaves free.  The collective has to finish inside the
        # ~180 us dgrad window: measured on B200 boxes (profiles/n8_probe.py, step time in ms, 5.6 MB bucket)
        #   2 ranks:  4 -> 0.928   8 -> 0.842   16 -> 0.846   32 -> 0.890   (no all-reduce 0.851, all-reduce afterwards 0.885)
        #   4 ranks:  4 -> 0.989   8 -> 0.885   16 -> 0.857   32 -> 0.868   (no all-reduce 0.852, all-reduce afterwards 0.892)
        #   8 ranks:  4 -> 1.277   8 -> 1.068   16 -> 0.971   32 -> 0.867   (no all-reduce 0.854, all-reduce afterwards 0.944)
        convs = list(head.conv2d_list)
        # bucket order: the R weights, then the R biases as one [R, C] block.  In symmetric memory when the box allows it: the
        # all-reduce is then our own peer-memory kernel (NVSwitch multimem reduce + broadcast, csrc/p2p_allreduce.cu).  Measured
        # (profiles/p2p_probe.py, 5.6 MB): 8 ranks 38 us on FOUR CTAs (NCCL: 178 / 104 / 57 us at 8 / 16 / 32 CTAs), 2 ranks 52 / 42 us
        # on 4 / 8 CTAs -- so the data-gradient GEMM cedes 4 SMs instead of the 8 / 16 / 32 an NCCL ring needs.
        super().__init__([m.weight for m in convs] + [m.bias for m in convs], symmetric=True, group=group)
        if overlap_ctas is None and os.environ.get("B200SEG_OVERLAP_CTAS"):
            overlap_ctas = int(os.environ["B200SEG_OVERLAP_CTAS"])           # (experiments)
        if overlap_ctas is None:
            world = dist.get_world_size() if is_distributed() else 1
            if self._symm is not None:
                # multimem path (> 2 ranks): 4 CTAs (38 us for the 5.6 MB bucket at 8 ranks).  Fewer looked as good in a long-running
                # probe at sustained clocks (profiles/dp_overlap_probe.py) but lose in the bench regime: 8 ranks, 20 timed steps,
                # 4 CTAs 0.767 ms per step, 2 CTAs 0.865 ms (the collective no longer fits under the data-gradient GEMM)
                # (two ranks, bench regime: 4 CTAs 0.748 / 0.751 ms per step, 8 CTAs 0.766 / 0.756 -- the multicast address exists at
                #  two ranks as well)
                overlap_ctas = 4
            else:
                overlap_ctas = 8 if world <= 2 else (16 if world <= 4 else 32)
        self.p2p_blocks = int(overlap_ctas) if overlap_ctas > 0 else 8
        self.R = len(convs)
        # The persistent dgrad GEMM fills every SM's shared memory, so nothing else can be resident beside it: it leaves
        # `overlap_ctas` SMs free, and the all-reduce runs on a dedicated NCCL communicator capped at that many CTAs.
        if self._symm is None and group is None and is_distributed() and dist.get_backend() == "nccl" and overlap_ctas > 0:
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(overlap_ctas)
                opts.config.min_ctas = 1
                group = dist.new_group(backend="nccl", pg_options=opts)
            except Exception:                        # older torch / NCCL without per-communicator CTA limits
                group = None
        if overlap_ctas > 0 and self.flat.is_cuda:
            from . import _lib
            _lib.gemm_set_overlap_sms(overlap_ctas)
        self.group = group
        self.stream = torch.cuda.Stream(device=self.flat.device)
        self.ready_event = torch.cuda.Event()
        self.ready_event.record()                    # materialise the CUDA event: its handle is passed through the C ABI
        self._pending = False
        nb = sum(m.bias.numel() for m in convs)
        self._bias_block = self.flat[self.flat.numel() - nb:].view(self.R, -1)
        # NCCL builds a communicator's channels lazily on its first collectives (tens of ms): do that here, on the (all-zero)
        # bucket, instead of inside the first training steps
        if is_distributed() and self.flat.is_cuda:
            with torch.cuda.stream(self.stream):
                for _ in range(3):
                    if self._symm is not None:
                        self.p2p_allreduce_mean_(self.p2p_blocks)
                    else:
                        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            torch.cuda.current_stream().wait_stream(self.stream)

    def weight_buffers(self):
        return self.views[:self.R]

    def bias_buffers(self):
        return self.views[self.R:2 * self.R]

    def set_bias_grads_(self, bias_grad: torch.Tensor):
        self._bias_block.copy_(bias_grad.unsqueeze(0).expand_as(self._bias_block))     # d/d b_r is the same for every branch

    def begin_allreduce_(self):
        for p, v in zip(self.params, self.views):
            p.grad = v
        if not is_distributed():
            return
        self.stream.wait_event(self.ready_event)     # weight gradients (and the bias block enqueued before them) are complete
        with torch.cuda.stream(self.stream):
            if self._symm is not None:
                self.p2p_allreduce_mean_(self.p2p_blocks)
            elif dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(dist.get_world_size(self.group))
        self._pending = True

    def step(self, optimizer):
        """``optimizer.step()`` for the bucket's parameters, enqueued on the bucket's stream directly behind the all-reduce
        (SURVEY 8f rank 4): with ``FusedSGD`` that is one K8 launch, so all-reduce + update both run underneath the
        data-gradient GEMM and the host never waits for the collective.  Safe while that GEMM is in flight: it reads the
        PACKED bf16 weights of this step's forward, not the parameters.  ``wait()`` (before the next forward) joins."""
        if not self._pending:
            return optimizer.step()
        with torch.cuda.stream(self.stream):
            return optimizer.step()

    def wait(self):
        if self._pending:
            torch.cuda.current_stream().wait_stream(self.stream)
            self._pending = False


def allreduce_mean_grads_(module: torch.nn.Module, group=None, bucket: Optional[FlatGradBucket] = None) -> FlatGradBucket:
    bucket = bucket or FlatGradBucket(module.parameters())
    bucket.allreduce_mean_(group)
    return bucket


def allreduce_confusion_(cm: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the int64 confusion matrix."""
    if is_distributed():
        dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=group)
    return cm


def allgather_frame_areas(local: torch.Tensor, n_frames: int, group=None) -> torch.Tensor:
    """Per-frame (intersection, union, target, output) areas of ALL frames in frame order, for the macro metrics of
    ``AverageMeter`` (utility.py:24-72: mean over frames of per-frame IoU / F1), when frames are sharded ``rank::world``
    (``shard_indices``).  ``local``: integer tensor [n_local, 4, C], row k = this rank's k-th frame (frame ``rank + k * world``).
    Returns [n_frames, 4, C] on every rank; integer areas, so the result is exact and order-independent."""
    if not is_distributed():
        return local[:n_frames]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per_rank = (n_frames + world - 1) // world                      # the largest shard; shorter shards are zero-padded
    padded = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = torch.empty((n_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, part in enumerate(parts):
        idx = shard_indices(n_frames, r, world)
        out[idx] = part[:len(idx)]
    return out
