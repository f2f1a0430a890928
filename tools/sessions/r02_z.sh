#!/bin/bash
# round 2, session z (1 GPU): choose_segment() with and without its
# fill-one-wave-first clause on windows the tuner does not measure (under 2^22
# cells, or SODA_CUDA_AUTOTUNE=0), GPU tests with the new rule
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
R="timeout 300 python tools/run_one.py"
{
for W in 0 1; do
export SODA_CUDA_WAVE_RULE=$W
$R jacobi2d 4096,960 --iterate 66 --tb 6 --reps 20
$R jacobi2d 8192,448 --iterate 66 --tb 6 --reps 20
$R jacobi2d 2048,1900 --iterate 66 --tb 6 --reps 20
$R blur 2000,1000 --iterate 2 --reps 20
$R sobel2d 4096,1000 --reps 20
$R jacobi3d 256,256,60 --iterate 32 --tb 2 --reps 20
SODA_CUDA_AUTOTUNE=0 $R jacobi2d 16384,640 --iterate 66 --tb 6 --reps 20
SODA_CUDA_AUTOTUNE=0 $R jacobi2d 16384,1152 --iterate 66 --tb 6 --reps 20
SODA_CUDA_AUTOTUNE=0 $R jacobi2d 16384,16384 --iterate 66 --tb 6 --reps 5
done
} > $O/r02z_wave_rule.jsonl 2> $O/r02z_wave_rule.err
python - <<PY
import json
rows = [json.loads(l) for l in open('$O/r02z_wave_rule.jsonl')]
half = len(rows) // 2
for a, b in zip(rows[:half], rows[half:]):
  print(a['program'], a['extent'], 'tb', a['tb'], 'ms/pass without %.4f with %.4f  (%+.0f %%)' % (a['ms_per_pass'], b['ms_per_pass'], 100 * (a['ms_per_pass'] / b['ms_per_pass'] - 1)))
PY
tail -3 $O/r02z_wave_rule.err
unset SODA_CUDA_WAVE_RULE
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02z_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02z_pytest_gpu.log
