#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
export SODA_CUDA_VERBOSE=1
python -m pytest tests -m gpu -x -q > $O/pytest_gpu6.log 2>&1; tail -3 $O/pytest_gpu6.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1d.json 2> $O/bench_r1d.err; cat $O/bench_r1d.json; grep soda_cuda: $O/bench_r1d.err | head
python tools/bench_configs.py run > $O/configs_v3.log 2> $O/configs_v3.err; cat $O/configs_v3.log; grep soda_cuda: $O/configs_v3.err
