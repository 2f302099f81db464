"""torchrun probe at N ranks: step time without all-reduce, with the bucket all-reduce overlapped under the dgrad GEMM for several
NCCL CTA caps (= SMs the dgrad GEMM leaves free), with the all-reduce issued after backward on the default communicator, and the
all-reduce alone.   torchrun --nproc-per-node N profiles/n8_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth, distributed as D

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
b200.set_feature_pack_cache(0)


def timeit(fn, iters=20):
    for _ in range(5):
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


def step(bucket, mode):
    xg = x.detach().requires_grad_(True)
    for p in head.parameters():
        p.grad = None
    loss, _ = head.forward_loss(xg, labels, grad_bucket=bucket if mode == "overlap" else None)
    loss.backward()
    if mode == "overlap":
        bucket.wait()
    elif mode == "after":
        bucket.allreduce_mean_()


res = {}
res["no all-reduce"] = timeit(lambda: step(None, "plain"))
plain_bucket = D.FlatGradBucket(head.parameters())
res["all-reduce after backward (default comm)"] = timeit(lambda: step(plain_bucket, "after"))
res["all-reduce alone, default comm (us)"] = timeit(lambda: dist.all_reduce(plain_bucket.flat, op=dist.ReduceOp.AVG)) * 1e3
for ctas in (4, 8, 16, 32):
    bucket = D.HeadGradBucket(head, overlap_ctas=ctas)
    res[f"overlapped, {ctas} CTAs"] = timeit(lambda: step(bucket, "overlap"))
    res[f"all-reduce alone, {ctas}-CTA comm (us)"] = timeit(lambda: dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG, group=bucket.group)) * 1e3
if rank == 0:
    print(f"world={world}")
    for k, v in res.items():
        print(f"  {k}: {v:.3f}")
dist.destroy_process_group()
