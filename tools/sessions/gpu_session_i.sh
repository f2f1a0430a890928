#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu10.log 2>&1; tail -5 $O/pytest_gpu10.log
python tools/bench_layout.py > $O/layout_bench3.log 2>&1; cat $O/layout_bench3.log
# the bench kernel with the segment the measurement chose (335 rows), no tuning launches:
# launches = 11 (warm-up) + 22; the 13th soda launch is a time-block-6 pass of the timed region
SODA_CUDA_AUTOTUNE=0 SODA_CUDA_SEGMENT=335 ncu --set full --clock-control none --import-source on -k regex:soda_stream2d --launch-skip 12 --launch-count 1 -o $O/prof_j2d_tb6 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full_tb6.log 2>&1
tail -3 $O/ncu_full_tb6.log
ncu --set full --clock-control none --import-source on -k regex:unpack_kernel -c 1 -o $O/prof_layout_unpack -f python tools/bench_layout.py > $O/ncu_layout3.log 2>&1
