#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu14.log 2>&1; tail -5 $O/pytest_gpu14.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1g.json 2> $O/bench_r1g.err; cat $O/bench_r1g.json
python tools/bench_configs.py run > $O/configs_v8.log 2> $O/configs_v8.err; cat $O/configs_v8.log
python tools/bench_layout.py > $O/layout_bench4.log 2>&1; cat $O/layout_bench4.log
