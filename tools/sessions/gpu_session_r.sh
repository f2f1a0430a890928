#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu18.log 2>&1; tail -4 $O/pytest_gpu18.log
python tools/bench_layout.py > $O/layout_bench6.log 2>&1; cat $O/layout_bench6.log
true
