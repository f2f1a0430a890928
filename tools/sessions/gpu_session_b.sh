#!/bin/bash
# GPU session: parity tests, 3-D tuning sweep, ncu of the 2-D time-block-8 kernel
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu4.log 2>&1; tail -3 $O/pytest_gpu4.log
python tools/tune3d.py run > $O/tune3d_v1.log 2>&1
tail -5 $O/tune3d_v1.log
NCU="ncu --set full --clock-control none --import-source on"
python tools/run_one.py jacobi2d 16384,16384 --iterate 16 --tb 8 > $O/j2d_tb8.json 2>&1 && \
$NCU -k regex:soda_stream2d -s 2 -c 1 -o $O/prof_j2d_tb8 -f python tools/run_one.py jacobi2d 16384,16384 --iterate 16 --tb 8 --reps 1 --warmup 1 > $O/ncu_j2d_tb8.log 2>&1
cat $O/j2d_tb8.json
