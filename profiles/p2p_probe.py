"""Correctness and timing of the peer-memory mean all-reduce kernel (csrc/p2p_allreduce.cu) against NCCL, run under torchrun."""
import os
import sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnd_semantic_segmentation_b200 import _lib, distributed as D

rank, world, local = D.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
_lib.load()
for numel in (1_400_908, 5_057_702):
    buf, hdl = D.symmetric_flat_buffer(numel, dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(buf.numel(), device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    for mc in (True, False):
        for blocks in (4, 8, 16, 32):
            buf.copy_(src)
            torch.cuda.synchronize(); dist.barrier()
            _lib.p2p_allreduce_mean(hdl.buffer_ptrs, (hdl.multicast_ptr or 0) if mc else 0, hdl.signal_pad_ptrs, hdl.rank, hdl.world_size,
                                    buf.numel(), blocks=blocks, device=dev)
            torch.cuda.synchronize()
            err = ((buf - ref).abs().max() / ref.abs().max()).item()
            allb = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(allb, buf)
            same = all(torch.equal(allb[0], t) for t in allb)
            # timing: 20 back-to-back calls
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                _lib.p2p_allreduce_mean(hdl.buffer_ptrs, (hdl.multicast_ptr or 0) if mc else 0, hdl.signal_pad_ptrs, hdl.rank, hdl.world_size,
                                        buf.numel(), blocks=blocks, device=dev)
            e1.record(); torch.cuda.synchronize()
            if rank == 0:
                print(f"world {world} numel {numel} multimem={mc} blocks={blocks}: rel err vs NCCL AVG {err:.2e}, identical on all ranks {same}, {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
    t = src.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"world {world} numel {numel} NCCL all_reduce AVG (default CTAs): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
