#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu9.log 2>&1; tail -5 $O/pytest_gpu9.log
python tools/bench_layout.py > $O/layout_bench2.log 2>&1; cat $O/layout_bench2.log
ncu --set full --clock-control none --import-source on -k regex:soda_stream2d --launch-skip 118 --launch-count 2 -o $O/prof_j2d_tb6 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full_tb6.log 2>&1
tail -3 $O/ncu_full_tb6.log
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -c 1 -o $O/prof_layout2 -f python tools/bench_layout.py > $O/ncu_layout2.log 2>&1
ls -la $O/*.ncu-rep | tail -4
