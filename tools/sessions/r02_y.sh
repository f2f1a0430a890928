#!/bin/bash
# round 2, session y (1 GPU): timeline of the host pipeline (the pipeline's own
# events with timing) for equal and ramped chunk layouts
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
SODA_CUDA_PIPELINE_TRACE=1 timeout 300 python tools/e2e_ab.py 16 -16 24 -24 32 > $O/r02y_e2e_traced.jsonl 2> $O/r02y_trace.err
cat $O/r02y_e2e_traced.jsonl
python tools/pipeline_trace_summary.py $O/r02y_trace.err --chunks > $O/r02y_trace_summary.txt; cat $O/r02y_trace_summary.txt
timeout 300 python tools/e2e_ab.py 16 -16 16 -16 > $O/r02y_e2e.jsonl 2> $O/r02y_e2e.err; cat $O/r02y_e2e.jsonl
