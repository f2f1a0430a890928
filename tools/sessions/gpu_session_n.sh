#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu15.log 2>&1; tail -5 $O/pytest_gpu15.log
python tools/bench_layout.py > $O/layout_bench5.log 2>&1; cat $O/layout_bench5.log
python tools/tune2d.py run > $O/tune2d_v5.log 2>&1
SODA_TUNE_SET=final python tools/tune3d.py run > $O/tune3d_v3.log 2>&1
tail -2 $O/tune2d_v5.log $O/tune3d_v3.log
