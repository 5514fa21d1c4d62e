"""Per-iteration and fixed cost of the Sinkhorn kernels: time vs number of iterations, K = 200.
  single  : one resident call (all SMs)                      B rows
  dual    : timet_sinkhorn_pair, two problems side by side   2 x B rows, each on half of the SMs
Run on the GPU box:  python profiles/sinkhorn_iters.py > gpurun_out/sinkhorn_iters.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import timetuning_b200 as tb
from timetuning_b200 import synth


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for clips in (32, 16, 8):
    B = clips * 784
    s0 = torch.from_numpy(synth.cosine_scores(B, 200, seed=4)).cuda()
    s1 = torch.from_numpy(synth.cosine_scores(B, 200, seed=5)).cuda()
    o0, o1 = torch.empty_like(s0).view(clips, 784, 200), torch.empty_like(s0).view(clips, 784, 200)
    print(f"B = {B} rows ({clips} clips): single mode {tb.ops.sinkhorn_mode(B, 200)}, pair mode {tb.ops.sinkhorn_pair_mode(B, 200)}")
    res = {}
    for iters in (1, 2, 5, 10, 20):
        a = timed(lambda: tb.sinkhorn_from_scores(s0, 0.05, iters, out=o0))
        b = timed(lambda: tb.sinkhorn_pair_from_scores(s0, s1, 0.05, iters, out0=o0, out1=o1))
        res[iters] = (a, b)
        print(f"  iters={iters:2d}: single {a:7.1f} us   pair {b:7.1f} us (2 problems)")
    print(f"  per iteration: single {(res[20][0] - res[10][0]) / 10:.2f} us, pair {(res[20][1] - res[10][1]) / 10:.2f} us; "
          f"1 iteration: single {res[1][0]:.1f} us, pair {res[1][1]:.1f} us")
