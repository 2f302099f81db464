"""torchrun probe (N >= 2): data-parallel head training with the gradient bucket's all-reduce AND the FusedSGD step enqueued
behind it on the bucket's stream (HeadGradBucket.step), both underneath the dgrad GEMM.
Checks: every rank ends with bit-identical parameters; they equal a single-process run on the mean of the ranks' gradients
(computed here by all-gathering the per-rank gradients) to 1e-6.  Times: step with the optimizer on the bucket stream vs
bucket.wait() followed by optimizer.step() on the compute stream."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth, distributed as D

RATES = [6, 12, 18, 24]
rank, world, local = D.env_rank_world()
torch.cuda.set_device(local)
D.init_from_env("nccl")
dev = torch.device("cuda", local)
n, cin, h, w, H, W, C = synth.WORKLOADS["train_b8_512x1024"]
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)          # same seed: identical initial weights on every rank
shadow = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
shadow.load_state_dict(head.state_dict())
xs = [synth.make_features(n, cin, h, w, seed=100 + 10 * k + rank, device=dev) for k in range(3)]      # different data per rank
ys = [synth.make_labels(n, H, W, C, seed=200 + 10 * k + rank, device=dev) for k in range(3)]
bucket = D.HeadGradBucket(head)
hyper = dict(lr=2.5e-3, momentum=0.9, weight_decay=5e-4)
opt = b200.FusedSGD(head.parameters(), **hyper)
opt_shadow = torch.optim.SGD(shadow.parameters(), foreach=False, **hyper)

for k in range(3):
    # data-parallel step: weight gradients -> bucket -> all-reduce (mean) -> FusedSGD, all behind the wgrad GEMM
    xg = xs[k].detach().requires_grad_(True)
    loss, _ = head.forward_loss(xg, ys[k], grad_bucket=bucket)
    loss.backward()
    bucket.step(opt)
    bucket.wait()
    # shadow: local gradients without the bucket, mean over ranks by all_gather, torch.optim on the result
    for p in shadow.parameters():
        p.grad = None
    xg = xs[k].detach().requires_grad_(True)
    loss_s, _ = shadow.forward_loss(xg, ys[k])
    loss_s.backward()
    for p in shadow.parameters():
        parts = [torch.empty_like(p.grad) for _ in range(world)]
        dist.all_gather(parts, p.grad.contiguous())
        p.grad = torch.stack(parts).double().mean(0).float()
    opt_shadow.step()

torch.cuda.synchronize()
flat = torch.cat([p.detach().reshape(-1) for p in head.parameters()])
parts = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(parts, flat)
identical = all(torch.equal(parts[0], q) for q in parts[1:])
flat_s = torch.cat([p.detach().reshape(-1) for p in shadow.parameters()])
init = torch.cat([p.detach().reshape(-1) for p in b200.ASPP_Classifier_V2(cin, RATES, RATES, C).state_dict().values()]) if False else None
err = ((flat.double() - flat_s.double()).abs().max() / flat_s.double().abs().max()).item()
torch.manual_seed(0)
w0 = torch.cat([p.detach().reshape(-1) for p in b200.ASPP_Classifier_V2(cin, RATES, RATES, C).parameters()]).to(dev)
upd, upd_s = flat.double() - w0.double(), flat_s.double() - w0.double()
err_update = ((upd - upd_s).abs().max() / upd_s.abs().max()).item()


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def step(mode):
    xg = xs[0].detach().requires_grad_(True)
    loss, _ = head.forward_loss(xg, ys[0], grad_bucket=bucket)
    loss.backward()
    if mode == "bucket_stream":
        bucket.step(opt)
        bucket.wait()
    elif mode == "after_wait":
        bucket.wait()
        opt.step()
    else:
        bucket.wait()


t_none = timeit(lambda: step("none"))
t_after = timeit(lambda: step("after_wait"))
t_bucket = timeit(lambda: step("bucket_stream"))
if rank == 0:
    print(f"world={world}: params identical across ranks: {identical}; vs single-process mean-gradient torch.optim.SGD: "
          f"params rel err {err:.2e}, accumulated update rel err {err_update:.2e}")
    print(f"step (fwd + bwd + overlapped all-reduce) {t_none:.3f} ms; + FusedSGD after wait {t_after:.3f} ms; "
          f"+ FusedSGD on the bucket stream behind the all-reduce {t_bucket:.3f} ms")
    assert identical and err <= 1e-6 and err_update <= 1e-3, (identical, err, err_update)
dist.destroy_process_group()
