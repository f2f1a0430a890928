#!/bin/bash
# round 2, session c (1 GPU): everything after the slab ABI, the planner's
# 8-cell / time-block-8 choice, math precision, the re-measured segment tuning
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $O/r02c_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02c_bench.json 2> $O/r02c_bench.err; echo "bench exit $?"
tail -c 1200 $O/r02c_bench.json; tail -5 $O/r02c_bench.err
timeout 900 python tools/tb_sweep.py run > $O/r02c_tb_sweep.jsonl 2> $O/r02c_tb_sweep.err
cut -c1-200 $O/r02c_tb_sweep.jsonl
timeout 600 ncu -k regex:soda --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02c_bench_launches.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/r02c_ncu_list.log 2>&1; echo "ncu list exit $?"
SODA_CUDA_AUTOTUNE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:soda_stream2d --launch-skip 10 --launch-count 1 -o $O/r02c_prof_j2d python bench.py --steps 2 --warmup 1 --headline-only > $O/r02c_ncu_full.log 2>&1; echo "ncu full exit $?"
