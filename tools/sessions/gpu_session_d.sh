#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu5.log 2>&1; tail -3 $O/pytest_gpu5.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1c.json 2> $O/bench_r1c.err; cat $O/bench_r1c.json
python tools/bench_configs.py run > $O/configs_v2.log 2>&1; cat $O/configs_v2.log
