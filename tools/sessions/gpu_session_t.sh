#!/bin/bash
# ncu --set full of the shipped 3-D kernels (planner defaults)
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
export SODA_CUDA_AUTOTUNE=0
for spec in "heat3d 2" "jacobi3d 2" "jacobi3d 1"; do
  set -- $spec
  $NCU -k regex:soda_stream3d -s 1 -c 1 -o $O/prof_$1_tb$2_final -f python tools/run_one.py $1 512,512,512 --iterate 4 --tb $2 --reps 1 --warmup 1 > $O/ncu_$1_tb$2_final.log 2>&1
done
$NCU -k regex:soda_stream3d -s 1 -c 1 -o $O/prof_denoise3d_final -f python tools/run_one.py denoise3d 512,512,512 --reps 1 --warmup 1 > $O/ncu_denoise3d_final.log 2>&1
ls -la $O/*_final.ncu-rep
