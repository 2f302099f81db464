"""cProfile of the host side of one small train step (per-GPU batch 1 of deeplabv2_r101_tgt_self_distill): where the ~0.5 ms of
CPU time per eager step goes.   python profiles/cpu_profile_step.py"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rnd_semantic_segmentation_b200 as b200
from rnd_semantic_segmentation_b200 import synth

RATES = [6, 12, 18, 24]
n, cin, h, w, H, W, C = synth.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "deeplabv2_r101_tgt_self_distill"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head = b200.ASPP_Classifier_V2(cin, RATES, RATES, C).to(dev)
x = synth.make_features(n, cin, h, w, device=dev)
labels = synth.make_labels(n, H, W, C, device=dev)
b200.set_feature_pack_cache(0)
from rnd_semantic_segmentation_b200 import _lib
_lib.set_step_graphs(os.environ.get("GRAPHS", "1") != "0")
params = list(head.parameters())


def step():
    xg = x.detach().requires_grad_(True)
    for p in params:
        p.grad = None
    loss, _ = head.forward_loss(xg, labels)
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"CPU enqueue {1e3 * (t1 - t0) / 200:.3f} ms/step, wall incl. drain {1e3 * (t2 - t0) / 200:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
print("step graphs (replays, captures):", _lib.step_graph_stats())
