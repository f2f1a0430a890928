#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu16.log 2>&1; tail -3 $O/pytest_gpu16.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1h.json 2> $O/bench_r1h.err; cat $O/bench_r1h.json
python tools/bench_configs.py run > $O/configs_v9.log 2> $O/configs_v9.err; cat $O/configs_v9.log
SODA_TUNE_SET=final python tools/tune3d.py run denoise3d > $O/tune3d_v4.log 2>&1; cat $O/tune3d_v4.log
true
