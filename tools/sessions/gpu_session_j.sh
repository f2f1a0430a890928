#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu11.log 2>&1; tail -5 $O/pytest_gpu11.log
python tools/bench_configs.py run > $O/configs_v5.log 2> $O/configs_v5.err; cat $O/configs_v5.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1f.json 2> $O/bench_r1f.err; cat $O/bench_r1f.json
