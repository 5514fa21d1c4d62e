"""FF select stage alone at BASELINE configs[1] (for `ncu -k regex:ff_tc_persist --set full --import-source on`).
    ncu ... python profiles/prof_select.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import timetuning_b200 as tb
from timetuning_b200 import synth

bs, fs, sr, D = 32, 8, 28, 384
feats = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=1)).cuda()
plan = tb.FFPlan(bs, fs, sr, sr, D, 200, 7, 6, 5)
plan.prepare(feats)
for _ in range(3):
    plan.select(tb.FF_TC)
torch.cuda.synchronize()
print(plan.stats())
