#!/bin/bash
# 16-cell lanes (two TMA boxes per strip) against 8-cell lanes for 8/16-bit cells
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_half.py tests/test_params.py tests/test_gpu_parity.py -m gpu -x -q -k "half or narrow or blur or sobel or erosion or xcorr" > $O/pytest_lanes16.log 2>&1
tail -3 $O/pytest_lanes16.log
rm -f $O/lanes16.jsonl
run() { timeout 300 python tools/run_one.py "$@" --reps 5 >> $O/lanes16.jsonl 2>> $O/lanes16.err; }
for opts in '{}' '{"cells": 8}'; do
  run blur 16000,16384 --iterate 2 --tb 2 --options "$opts"
  run blur 16000,16384 --iterate 2 --tb 1 --options "$opts"
  run blur 2000,16384 --iterate 2 --tb 2 --options "$opts"
  run sobel2d 16384,16384 --options "$opts"
  run jacobi2d_half 16384,16384 --iterate 64 --tb 4 --options "$opts"
  run jacobi2d_half 16384,16384 --iterate 64 --tb 6 --options "$opts"
done
cat $O/lanes16.jsonl; tail -3 $O/lanes16.err
