"""Kernel-only time of the FF tensor-core nomination kernel at BASELINE configs[1] (select_timed events), per variant
(persistent kernel with its TIMET_TC_PFLAGS attribution switches, static schedule, per-tile kernel).  Also checks the
selection against the exact engine (meaningless for the variants that skip the scan).

    python profiles/tc_kernel_time.py [variant ...]"""
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
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); b.record()
for _ in range(3): plan.select_timed(tb.FF_TC, a, b)
torch.cuda.synchronize()
ks, ss = [], []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(10):
    e0.record(); plan.select_timed(tb.FF_TC, a, b); e1.record(); torch.cuda.synchronize()
    ks.append(a.elapsed_time(b)); ss.append(e0.elapsed_time(e1))
st = plan.stats()
sel_tc = [plan.selection(c, t) for c in (0, 31) for t in (1, 7)]
plan.select(tb.FF_EXACT)
sel_ex = [plan.selection(c, t) for c in (0, 31) for t in (1, 7)]
same = all(all(torch.equal(x, y) for x, y in zip(p, q)) for p, q in zip(sel_tc, sel_ex))
print("kernel ms %%.4f  select ms %%.4f  identical=%%s  %%s" %% (sum(ks) / len(ks), sum(ss) / len(ss), same, st))
''' % ROOT
VARIANTS = (("persistent", {}), ("raster-tiles", {"TIMET_TC_PFLAGS": "256"}), ("raster-noappend", {"TIMET_TC_PFLAGS": "258"}), ("wait-hint", {"TIMET_TC_PFLAGS": "512"}), ("no-wait-compaction", {"TIMET_TC_PFLAGS": "8192"}), ("two-stage-ring", {"TIMET_TC_PFLAGS": "16384"}), ("two-stage-ring-mma-only", {"TIMET_TC_PFLAGS": "16385"}), ("two-stage-ring-noappend", {"TIMET_TC_PFLAGS": "16386"}), ("opp-min-2", {"TIMET_TC_PFLAGS": str(2 << 20)}), ("opp-min-4", {"TIMET_TC_PFLAGS": str(4 << 20)}), ("opp-min-10", {"TIMET_TC_PFLAGS": str(10 << 20)}), ("opp-min-16", {"TIMET_TC_PFLAGS": str(16 << 20)}), ("top-down-rows", {"TIMET_TC_PFLAGS": "32768"}), ("two-stage-top-down", {"TIMET_TC_PFLAGS": str(16384 + 32768)}), ("no-final-merge", {"TIMET_TC_PFLAGS": "1024"}), ("no-inloop-compaction", {"TIMET_TC_PFLAGS": "2048"}), ("no-merge-no-compaction", {"TIMET_TC_PFLAGS": "3072"}), ("wait-sleep", {"TIMET_TC_PFLAGS": "32"}), ("thr-per-tile", {"TIMET_TC_PFLAGS": "64"}), ("static-schedule", {"TIMET_TC_DYN": "0"}), ("oldest-first", {"TIMET_TC_PFLAGS": "8"}),
            ("thr-per-tile", {"TIMET_TC_PFLAGS": "64"}), ("groups-own-buffers", {"TIMET_TC_PFLAGS": "128"}), ("four-tmem-buffers", {"TIMET_TC_NBUF": "4"}), ("mma-only", {"TIMET_TC_PFLAGS": "1"}),
            ("scan-noappend", {"TIMET_TC_PFLAGS": "2"}), ("tmem-loads-only", {"TIMET_TC_PFLAGS": "4"}),
            ("per-tile", {"TIMET_TC_PERSIST": "0"}), ("per-tile-mma-only", {"TIMET_TC_PERSIST": "0", "TIMET_TC_FLAGS": "1"}),
            ("per-tile-noappend", {"TIMET_TC_PERSIST": "0", "TIMET_TC_FLAGS": "2"}),
            ("stages-2", {"TIMET_TC_STAGES": "2"}), ("mma-only-stages-2", {"TIMET_TC_PFLAGS": "1", "TIMET_TC_STAGES": "2"}))
for name, extra in VARIANTS:
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    out = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **extra), capture_output=True, text=True)
    print(f"{name}:", out.stdout.strip(), out.stderr.strip()[-600:])
