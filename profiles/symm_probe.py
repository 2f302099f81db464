"""Probe: torch symmetric memory rendezvous on this box (peer buffer / signal-pad pointers for the P2P all-reduce kernel)."""
import os
import sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local))
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "rank/world", hdl.rank, hdl.world_size, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs],
      "signal_pad_size", getattr(hdl, "signal_pad_size", None), "multicast_ptr", hex(getattr(hdl, "multicast_ptr", 0) or 0), flush=True)
t.fill_(rank + 1.0)
dist.barrier()
torch.cuda.synchronize()
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
print(rank, "peer value", peer[:2].tolist(), flush=True)
out = torch.ops.symm_mem.one_shot_all_reduce(t, "sum", dist.group.WORLD.group_name)
print(rank, "one_shot_all_reduce", out[:2].tolist(), flush=True)
dist.barrier()
dist.destroy_process_group()
