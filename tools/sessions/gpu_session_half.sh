#!/bin/bash
# `half` programs: parity on the GPU, then device-resident throughput of the
# binary16 jacobi2d (pairs and scalar cells)
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_half.py -m gpu -x -q > $O/pytest_half.log 2>&1
tail -5 $O/pytest_half.log
rm -f $O/half_perf.jsonl
for tb in 4 6 8 12; do
  timeout 300 python tools/run_one.py jacobi2d_half 16384,16384 --iterate 64 --tb $tb --reps 3 >> $O/half_perf.jsonl 2>> $O/half_perf.err
done
timeout 300 python tools/run_one.py jacobi2d_half 16384,16384 --iterate 64 --tb 4 --options '{"no_pack": true}' --reps 3 >> $O/half_perf.jsonl 2>> $O/half_perf.err
cat $O/half_perf.jsonl; tail -3 $O/half_perf.err
