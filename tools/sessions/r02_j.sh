#!/bin/bash
# round 2, session j (1 GPU): how much would removing the dimension-0 halo
# redundancy buy in 3-D?  The same kernels on a 488-wide grid (four 128-cell
# strips with 120 valid cells fit exactly) against the 512-wide one (five).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
R="timeout 300 python tools/run_one.py"
{
for w in 488 512 608; do
  $R heat3d $w,512,512 --iterate 32 --tb 2
  $R jacobi3d $w,512,512 --iterate 32 --tb 2
  $R denoise3d $w,512,512
  for tb in 3 4; do
    $R jacobi3d $w,512,512 --iterate $((tb*4)) --tb $tb
    $R heat3d $w,512,512 --iterate $((tb*4)) --tb $tb
  done
done
} > $O/r02j_fit.jsonl 2> $O/r02j_fit.err
cut -c1-230 $O/r02j_fit.jsonl; tail -3 $O/r02j_fit.err
