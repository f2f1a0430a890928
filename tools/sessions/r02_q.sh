#!/bin/bash
# round 2, session q (1 GPU): programmatic dependent launch on / off
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02q_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02q_pytest_gpu.log
R="timeout 300 python tools/run_one.py"
{
for pdl in 1 0 1 0; do
  echo "{\"pdl\": $pdl}"
  SODA_CUDA_PDL=$pdl timeout 300 python bench.py --steps 30 --warmup 3 --headline-only 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({'bench_value': d['value'], 'ms_per_step': d['ms_per_step'], 'frac': d['roofline']['frac']}))"
  SODA_CUDA_PDL=$pdl $R heat3d 512,512,512 --iterate 32 --tb 2 --reps 5
  SODA_CUDA_PDL=$pdl $R jacobi3d 512,512,512 --iterate 32 --tb 2 --reps 5
  SODA_CUDA_PDL=$pdl $R jacobi3d 512,512,512 --iterate 32 --tb 1 --reps 5
  SODA_CUDA_PDL=$pdl $R blur 2000,16384 --iterate 2 --tb 1 --reps 20
  SODA_CUDA_PDL=$pdl $R seidel2d 16384,16384 --iterate 16 --reps 3
done
} > $O/r02q_pdl.jsonl 2> $O/r02q_pdl.err
cut -c1-250 $O/r02q_pdl.jsonl; tail -3 $O/r02q_pdl.err
