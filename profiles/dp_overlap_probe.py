"""torchrun probe at N ranks: where the data-parallel step's extra time over the single-rank step goes.
   plain step (no bucket) / bucket path with the all-reduce kernel skipped (ceded SMs + stream join only) / overlapped peer-memory
   all-reduce at several CTA counts / the same with the data-gradient GEMM ceding no SMs.
   torchrun --nproc-per-node N profiles/dp_overlap_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth, distributed as D, _lib

RATES = [6, 12, 18, 24]
rank, world, local = D.env_rank_world()
torch.cuda.set_device(local)
D.init_from_env("nccl")
dev = torch.device("cuda", local)
n, cin, h, w, H, W, C = synth.WORKLOADS["train_b8_512x1024"]
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev)
labels = synth.make_labels(n, H, W, C, device=dev)
params = list(head.parameters())


def timeit(fn, iters=40):
    for _ in range(8):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def step(bucket):
    xg = x.detach().requires_grad_(True)
    for p in params:
        p.grad = None
    loss, _ = head.forward_loss(xg, labels, grad_bucket=bucket)
    loss.backward()
    if bucket is not None:
        bucket.wait()


res = {}
_lib.set_step_graphs(True)
buckets = {c: D.HeadGradBucket(head, overlap_ctas=c) for c in (1, 2, 3, 4, 8)}
for rep in range(3):                                       # interleaved repeats: order / clock drift shows up as spread
    res[f"rep{rep}: plain step, no bucket"] = timeit(lambda: step(None))
    for ctas, bucket in buckets.items():
        _lib.gemm_set_overlap_sms(ctas)
        res[f"rep{rep}: overlapped all-reduce, {ctas} CTAs"] = timeit(lambda: step(bucket))
_lib.set_step_graphs(False)
res["direct launches: plain step, no bucket"] = timeit(lambda: step(None))
for ctas in (2, 4):
    _lib.gemm_set_overlap_sms(ctas)
    res[f"direct launches: overlapped all-reduce, {ctas} CTAs"] = timeit(lambda: step(buckets[ctas]))
if rank == 0:
    print(f"world={world}")
    for k, v in res.items():
        print(f"  {k}: {v:.4f}")
dist.destroy_process_group()
