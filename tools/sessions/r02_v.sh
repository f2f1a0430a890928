#!/bin/bash
# round 2, session v (1 GPU): --cuda-pow2-fma (fused c * x + acc for
# power-of-two literals): GPU tests, heat3d at time blocks 1-4 with and without
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02v_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02v_pytest_gpu.log
R="timeout 300 python tools/run_one.py"
{
$R heat3d 512,512,512 --iterate 32 --tb 2 --reps 5
$R heat3d 512,512,512 --iterate 32 --tb 2 --reps 5 --options '{"pow2_fma": true}'
$R heat3d 512,512,512 --iterate 32 --tb 1 --reps 5 --options '{"pow2_fma": true}'
$R heat3d 512,512,512 --iterate 12 --tb 3 --reps 5 --options '{"pow2_fma": true}'
$R heat3d 512,512,512 --iterate 16 --tb 4 --reps 5 --options '{"pow2_fma": true}'
$R heat3d 512,512,512 --iterate 32 --tb 2 --reps 5 --options '{"pow2_fma": true}'
} > $O/r02v_pow2_fma.jsonl 2> $O/r02v_pow2_fma.err
python - <<PY
import json
for l in open('$O/r02v_pow2_fma.jsonl'):
  d=json.loads(l); print(d['program'], d['tb'], d['options'], 'ms/pass %.4f'%d['ms_per_pass'], 'Gcell/s %.0f'%d['gcell_per_s'], 'frac %.3f'%d['frac'])
PY
tail -3 $O/r02v_pow2_fma.err
