"""CPU enqueue time per train step (no synchronisation inside the loop) vs GPU time per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth, distributed as D

RATES = [6, 12, 18, 24]
n, cin, h, w, H, W, C = synth.WORKLOADS["train_b8_512x1024"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev)
labels = synth.make_labels(n, H, W, C, device=dev)
for use_bucket in (False, True):
    bucket = D.HeadGradBucket(head) if use_bucket else None

    def step():
        xg = x.detach().requires_grad_(True)
        for p in head.parameters():
            p.grad = None
        loss, _ = head.forward_loss(xg, labels, grad_bucket=bucket)
        loss.backward()
        if bucket is not None:
            bucket.wait()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"bucket={use_bucket}: CPU enqueue {1e3 * (t1 - t0) / 20:.3f} ms/step, GPU {e0.elapsed_time(e1) / 20:.3f} ms/step")
