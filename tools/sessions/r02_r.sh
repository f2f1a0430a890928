#!/bin/bash
# round 2, session r (1 GPU): 3-D kernels that skip the patch rows in a node's
# dimension-1 halo; GPU tests, then the 3-D configs
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02r_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02r_pytest_gpu.log
R="timeout 300 python tools/run_one.py"
{
for i in 1 2; do
$R heat3d 512,512,512 --iterate 32 --tb 2 --reps 5
$R jacobi3d 512,512,512 --iterate 32 --tb 2 --reps 5
$R jacobi3d 512,512,512 --iterate 32 --tb 1 --reps 5
$R heat3d 512,512,512 --iterate 32 --tb 1 --reps 5
$R denoise3d 512,512,512 --reps 5
$R jacobi3d 512,512,512 --iterate 12 --tb 3 --reps 5
$R heat3d 512,512,512 --iterate 12 --tb 3 --reps 5
$R jacobi3d 512,512,512 --iterate 16 --tb 4 --reps 5
$R heat3d 512,512,512 --iterate 16 --tb 4 --reps 5
done
$R jacobi3d 512,512,512 --iterate 12 --tb 3 --reps 5 --options '{"rows": 32, "cy": 4}'
$R jacobi3d 512,512,512 --iterate 16 --tb 4 --reps 5 --options '{"rows": 32, "cy": 4}'
$R heat3d 512,512,512 --iterate 12 --tb 3 --reps 5 --options '{"rows": 32, "cy": 4}'
} > $O/r02r_dead_rows.jsonl 2> $O/r02r_dead_rows.err
python - <<PY
import json
for l in open('$O/r02r_dead_rows.jsonl'):
  d=json.loads(l); print(d['program'], d['tb'], d['options'], 'ms/pass %.4f'%d['ms_per_pass'], 'Gcell/s %.0f'%d['gcell_per_s'], 'frac %.3f'%d['frac'])
PY
tail -3 $O/r02r_dead_rows.err
{
for p in heat3d jacobi3d; do
  $R $p 512,512,512 --iterate 32 --tb 2 --reps 5 --options '{"no_edge_roles": true}'
done
$R denoise3d 512,512,512 --reps 5 --options '{"no_edge_roles": true}'
} > $O/r02r_no_roles.jsonl 2>> $O/r02r_dead_rows.err
python - <<PY
import json
for l in open('$O/r02r_no_roles.jsonl'):
  d=json.loads(l); print(d['program'], d['tb'], d['options'], 'ms/pass %.4f'%d['ms_per_pass'], 'Gcell/s %.0f'%d['gcell_per_s'], 'frac %.3f'%d['frac'])
PY
