"""Per-iteration cost of the resident Sinkhorn kernel: time vs number of iterations at B = 25 088, K = 200."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import timetuning_b200 as tb
from timetuning_b200 import synth
scores = torch.from_numpy(synth.cosine_scores(25088, 200, seed=4)).cuda()
res = {}
for iters in (1, 2, 4, 10, 20):
    for _ in range(3):
        tb.sinkhorn_from_scores(scores, 0.05, iters)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tb.sinkhorn_from_scores(scores, 0.05, iters)
    e1.record(); torch.cuda.synchronize()
    res[iters] = e0.elapsed_time(e1) / 20 * 1e3
    print(f"iters={iters:2d}: {res[iters]:7.1f} us per call")
print(f"per iteration ~ {(res[20] - res[10]) / 10:.2f} us; fixed ~ {res[1]:.1f} us")
