#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
export SODA_CUDA_VERBOSE=1
python -m pytest tests -m gpu -x -q > $O/pytest_gpu17.log 2>&1; tail -3 $O/pytest_gpu17.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke3.log 2>&1; tail -1 $O/smoke3.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1i.json 2> $O/bench_r1i.err; cat $O/bench_r1i.json; grep soda_cuda: $O/bench_r1i.err | head -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_tb6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_list_tb6.log 2>&1
SEG=$(grep "variant 0 slices 16384" $O/bench_r1i.err | head -1 | sed -E 's/.*segment ([0-9]+).*/\1/')
echo "segment chosen by the measurement: $SEG"
SODA_CUDA_AUTOTUNE=0 SODA_CUDA_SEGMENT=${SEG:-335} ncu --set full --clock-control none --import-source on -k regex:soda_stream2d --launch-skip 12 --launch-count 1 -o $O/prof_j2d_tb6 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full_tb6.log 2>&1
tail -2 $O/ncu_full_tb6.log
echo "SEG=$SEG" > $O/prof_j2d_tb6.segment
