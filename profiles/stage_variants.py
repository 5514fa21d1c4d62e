"""Stage times of the bench step under kernel-variant switches (one fresh process per variant: the library reads the
environment once).  Run on the GPU box:  python profiles/stage_variants.py [--config N] > gpurun_out/variants.jsonl"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [{}, {"TIMET_GATHER_BATCH": "5"}, {"TIMET_GATHER_BATCH": "7"}, {"TIMET_SK_DUAL": "0"}, {"TIMET_SC_STAGES": "4"},
            {"TIMET_SK_STREAMING": "1"}, {"TIMET_GATHER_L1": "0"}, {"TIMET_FIN_STAGED": "0"}, {"TIMET_TC_PFLAGS": "256"}, {"TIMET_TC_PFLAGS": "8192"}]
extra = [a for a in sys.argv[1:] if not a.startswith("--only=")]
only = [a[7:].split(",") for a in sys.argv[1:] if a.startswith("--only=")]
if only:      # --only=TIMET_FIN_BLOCKED,... : the default plus the variants that set one of these switches
    VARIANTS = [v for v in VARIANTS if not v or any(k in only[0] for k in v)]
for env in VARIANTS:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-e2e",
                          "--no-cpu-baseline", *extra], env=dict(os.environ, **env), capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(json.dumps({"env": env, "clips_per_s": round(d["value"], 1), "ms_per_step": round(d["ms_per_step"], 4),
                          "stage_ms": {k: round(v, 4) for k, v in d["stage_ms"].items()}}), flush=True)
    except Exception as e:
        print(json.dumps({"env": env, "error": str(e), "stderr": out.stderr[-400:]}), flush=True)
