#!/bin/bash
# round 2, session k (1 GPU): blur C1 (2000 x 16384, iterate 2) against the
# segment length, with and without the measured choice
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
R="timeout 300 python tools/run_one.py"
{
$R blur 2000,16384 --iterate 2 --tb 2 --reps 20 --warmup 3
for seg in 8 12 16 24 32 48 64 96 128 192 256 512 1024; do
  SODA_CUDA_AUTOTUNE=0 SODA_CUDA_SEGMENT=$seg $R blur 2000,16384 --iterate 2 --tb 2 --reps 20 --warmup 3
done
for seg in 32 64 128 256; do
  SODA_CUDA_AUTOTUNE=0 SODA_CUDA_SEGMENT=$seg $R blur 2000,16384 --iterate 2 --tb 1 --reps 20 --warmup 3
done
} > $O/r02k_blur_segments.jsonl 2> $O/r02k_blur_segments.err
cut -c1-200 $O/r02k_blur_segments.jsonl; tail -3 $O/r02k_blur_segments.err
