#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
export SODA_CUDA_VERBOSE=1
python -m pytest tests -m gpu -x -q > $O/pytest_gpu7.log 2>&1; tail -3 $O/pytest_gpu7.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1e.json 2> $O/bench_r1e.err; cat $O/bench_r1e.json; grep soda_cuda: $O/bench_r1e.err | head
python tools/bench_configs.py run > $O/configs_v4.log 2> $O/configs_v4.err; cat $O/configs_v4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_tb6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_list_tb6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tb6 --launch-skip 120 --launch-count 2 -o $O/prof_j2d_tb6 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full_tb6.log 2>&1
ls -la $O/*.ncu-rep | tail -3
