#!/bin/bash
# Round-end evidence on one B200 (run through gpurun):  bash profiles/final_artifacts.sh [quick]  -> gpurun_out/r2f_*
# 1 GPU tests, 2 default bench line, 3 launch list (ncu time only), 4 ncu --set full of every kernel, 5 the other configs
# ("quick": only config 3, whose grid uses the column-blocked tiles).
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r2f_pytest_gpu.log)"
timeout 400 python bench.py > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $O/r2f_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2f_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:sk_|ff_|scores_' -s 24 -c 9 -f -o $O/prof_r2f_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_r2f_final.log 2>&1; echo "ncu full rc=$? $(grep -c Profiling $O/ncu_r2f_final.log) kernels"
CFGS="1 3 4 5"; [ "${1:-}" = quick ] && CFGS="3"
for c in $CFGS; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > $O/r2f_bench_cfg${c}_n1.json 2> $O/r2f_bench_cfg${c}_n1.err; echo "cfg $c rc=$? $(python -c "import json,sys; d=json.load(open('$O/r2f_bench_cfg${c}_n1.json')); print(round(d['value'],1), d['unit'], round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})" 2>&1 | tail -1)"
done
