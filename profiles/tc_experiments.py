"""Attribution experiments for the FF tensor-core kernel (debug flags via TIMET_TC_FLAGS):
times plan.select(FF_TC) at BASELINE configs[1] with parts of the epilogue disabled."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import timetuning_b200 as tb
from timetuning_b200 import synth
bs, fs, sr, D = 32, 8, 28, 384
feats = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=1)).cuda()
plan = tb.FFPlan(bs, fs, sr, sr, D, 200, 7, 6, 5)
plan.prepare(feats)
for _ in range(3): plan.select(tb.FF_TC)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.select(tb.FF_TC)
e1.record(); torch.cuda.synchronize()
print("select ms", e0.elapsed_time(e1) / 10, plan.stats())
''' % ROOT
for flags in sys.argv[1:] or ["0", "1", "2"]:
    env = dict(os.environ, TIMET_TC_FLAGS=flags)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"flags={flags}:", out.stdout.strip(), out.stderr.strip()[-300:])
