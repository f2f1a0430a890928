#!/bin/bash
# round 2, session ab (1 GPU): the automatic chunk layout of the host pipeline
# (20 chunks, each 8 % shorter than the one before it) against 16 equal chunks,
# alternating, and other ratios; GPU tests with the new layout
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r02ab_chunk_decay.jsonl
run() {
  local name="$1" decay="$2" chunks="$3"
  SODA_CUDA_CHUNK_DECAY="$decay" timeout 200 python tools/e2e_ab.py $chunks 2>> $O/r02ab.err | python -c "
import sys, json
for l in sys.stdin:
  d = json.loads(l); d['layout'] = '$name'; print(json.dumps(d))" >> $O/r02ab_chunk_decay.jsonl
}
for rep in 1 2 3; do
run "16 equal" 1 16
run "auto: 20 x0.92" 0.92 0
done
run "auto: 20 x0.90" 0.90 0
run "auto: 20 x0.94" 0.94 0
run "20 equal" 1 20
run "24 x0.92" 0.92 -24
run "28 x0.92" 0.92 -28
run "16 x0.92" 0.92 -16
run "24 x0.94" 0.94 -24
run "16 equal" 1 16
run "auto: 20 x0.92" 0.92 0
python - <<PY
import json
for l in open('$O/r02ab_chunk_decay.jsonl'):
  d = json.loads(l); print('%-18s best %.2f mean %.2f ms' % (d['layout'], d['ms_best'], d['ms_mean']))
PY
tail -3 $O/r02ab.err
SODA_CUDA_PIPELINE_TRACE=1 timeout 200 python tools/e2e_ab.py 0 2> $O/r02ab_trace.err > /dev/null
python tools/pipeline_trace_summary.py $O/r02ab_trace.err --chunks > $O/r02ab_trace_summary.txt; cat $O/r02ab_trace_summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02ab_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02ab_pytest_gpu.log
