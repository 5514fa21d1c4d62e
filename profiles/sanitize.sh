#!/bin/bash
# compute-sanitizer on the final kernels (small shapes: the tools slow kernels down 10-100x).  Run on the GPU box:
#   bash profiles/sanitize.sh   ->  gpurun_out/r2_sanitizer_{memcheck,racecheck,synccheck}.log
set -u
CS=/usr/local/cuda/bin/compute-sanitizer
TESTS="tests/test_gpu_ties.py tests/test_gpu_tc.py::test_tc_engine_is_bit_identical_to_exact tests/test_gpu_tc.py::test_tc_column_blocked_query_tiles tests/test_gpu_parity.py::test_timet_step_cfg1_golden tests/test_gpu_parity.py::test_sinkhorn_golden tests/test_gpu_parity.py::test_sinkhorn_strided_output_into_label_frames tests/test_gpu_parity.py::test_cosine_scores_multi_and_autograd tests/test_gpu_parity.py::test_non_square_grid_drop_ins tests/test_gpu_parity.py::test_eval_tail_upsample_argmax"
for tool in memcheck racecheck synccheck; do
  $CS --tool $tool --error-exitcode 86 --print-limit 20 python -m pytest $TESTS -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r2_sanitizer_$tool.log | tr '\n' ' ')"
done
# the persistent tcgen05 kernel + dual Sinkhorn at a mid-size shape (4 clips of configs[1])
$CS --tool memcheck --error-exitcode 86 python -c "
import torch, numpy as np
from timetuning_b200 import synth
from timetuning_b200.step import StepRunner
bs, fs, sr, D, K = 4, 8, 28, 384, 200
bb = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=1)).cuda()
hd = synth.head_features(bb[:, [0, -1]].cpu().numpy(), 256, seed=2)
pr = torch.from_numpy(synth.prototypes(K, 256, seed=3)).cuda()
r = StepRunner(bs, fs, sr, D, 256, K)
for _ in range(2): r.run(torch.from_numpy(np.ascontiguousarray(hd[:, 0])).cuda(), torch.from_numpy(np.ascontiguousarray(hd[:, 1])).cuda(), bb, pr)
torch.cuda.synchronize(); print('step ok', r.plan.stats())
" > gpurun_out/r2_sanitizer_memcheck_step.log 2>&1
echo "memcheck step rc=$? $(grep -E 'ERROR SUMMARY|step ok' gpurun_out/r2_sanitizer_memcheck_step.log | tr '\n' ' ')"
