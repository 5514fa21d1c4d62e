"""Per-CTA timeline of the FF tensor-core kernel (TIMET_TC_TRACE=1): prints, by number of key tiles, the mean
time between the 8 stamps: 0 entry, 1 setup done, 2 query tile loaded, 3 last MMA issued, 4 last tile scanned
(group 0), 5 lists exchanged, 6 published, 7 teardown."""
import ctypes as C
import os
import sys

os.environ["TIMET_TC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import timetuning_b200 as tb
from timetuning_b200 import _cabi, synth
from timetuning_b200.ops import _ptr, _stream

bs, fs, sr, D = 32, 8, 28, 384
feats = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=1)).cuda()
plan = tb.FFPlan(bs, fs, sr, sr, D, 200, 7, 6, 5)
plan.prepare(feats)
for _ in range(3):
    plan.select(tb.FF_TC)
torch.cuda.synchronize()
n = min(4096, bs * 7 * 7)
out = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
_cabi.check(_cabi.lib().timet_debug_tc_trace(C.byref(plan.params), _ptr(plan.workspace), plan.nbytes, _ptr(out), n, _stream()), "trace")
t = out.cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
print(f"kernel span {(t[:, 7].max() - t0) / 1e3:.1f} us over {n} CTAs; CTA lifetime mean {(t[:, 7] - t[:, 0]).mean() / 1e3:.1f} us")
d = np.diff(t, axis=1) / 1e3
life = (t[:, 7] - t[:, 0]) / 1e3
# group CTAs by lifetime quantiles (proxy for number of key tiles)
order = np.argsort(life)
for name, idx in (("shortest 20%", order[: n // 5]), ("middle 20%", order[2 * n // 5: 3 * n // 5]), ("longest 20%", order[-n // 5:])):
    m = d[idx].mean(axis=0)
    print(f"{name:13s} life {life[idx].mean():6.1f} us | setup {m[0]:5.1f} | A-load {m[1]:5.1f} | MMA issue {m[2]:6.1f} | "
          f"scan tail {m[3]:6.1f} | exchange {m[4]:5.1f} | publish {m[5]:5.1f} | teardown {m[6]:5.1f}")
