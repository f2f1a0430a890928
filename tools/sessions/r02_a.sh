#!/bin/bash
# round 2, session a: GPU tests (incl. BASELINE-size parity), bench line, time
# block sweep for the planner's model, host<->device copy peak, launch list
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/r02a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02a_pytest_gpu.log
tail -5 $O/r02a_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench exit $?"
tail -c 3000 $O/r02a_bench.json; tail -5 $O/r02a_bench.err
timeout 900 python tools/tb_sweep.py run > $O/r02a_tb_sweep.jsonl 2> $O/r02a_tb_sweep.err
cat $O/r02a_tb_sweep.jsonl | cut -c1-230
./tools/probe/pcie_probe 512 > $O/r02a_pcie_1gpu.jsonl 2>&1; cat $O/r02a_pcie_1gpu.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02a_bench_launches.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/r02a_ncu_bench.log 2>&1; echo "ncu exit $?"
