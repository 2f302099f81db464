"""torchrun probe: step time with / without the gradient all-reduce, and the all-reduce alone (5.6 MB, mean)."""
import os, sys, time
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
bucket = D.HeadGradBucket(head)


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
    return e0.elapsed_time(e1) / iters


def step(sync_mode):
    xg = x.detach().requires_grad_(True)
    for p in head.parameters():
        p.grad = None
    loss, _ = head.forward_loss(xg, labels, grad_bucket=bucket if sync_mode != "plain" else None)
    loss.backward()
    if sync_mode == "overlap":
        bucket.wait()
    elif sync_mode == "plain":
        pass


t_plain = timeit(lambda: step("plain"))
t_overlap = timeit(lambda: step("overlap"))
flat = bucket.flat
t_ar_default = timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
t_ar_capped = timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=bucket.group))
if rank == 0:
    print(f"world={world}: step without all-reduce {t_plain:.3f} ms, overlapped bucket step {t_overlap:.3f} ms, "
          f"all-reduce alone default comm {t_ar_default * 1e3:.1f} us, 8-CTA comm {t_ar_capped * 1e3:.1f} us")
dist.destroy_process_group()
