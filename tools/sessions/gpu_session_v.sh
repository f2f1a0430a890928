#!/bin/bash
# how much does the bench depend on the segment length, and does the measured choice find the best?
cd "$(dirname "$0")/../.."
O=gpurun_out
: > $O/bench_segments.log
for rep in 1 2; do
  for seg in 0 171 256 342 468 512 683 1024 1366; do
    if [ $seg = 0 ]; then
      v=$(SODA_CUDA_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2> $O/tmp.err | python -c "import sys,json; b=json.load(sys.stdin); print(b['value'], b['roofline']['ms_per_launch'])")
      echo "rep $rep tuned $(grep 'variant 0' $O/tmp.err | head -1 | sed -E 's/.*-> //') : $v" | tee -a $O/bench_segments.log
    else
      v=$(SODA_CUDA_AUTOTUNE=0 SODA_CUDA_SEGMENT=$seg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; b=json.load(sys.stdin); print(b['value'], b['roofline']['ms_per_launch'])")
      echo "rep $rep segment $seg : $v" | tee -a $O/bench_segments.log
    fi
  done
done
