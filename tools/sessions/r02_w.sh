#!/bin/bash
# round 2, session w (1 GPU): the chunk windows of the host pipeline get a
# measured segment length (tuner threshold 2^24 -> 2^22 cells).  e2e against the
# chunk count with the old and the new threshold, a per-chunk trace of the
# pipeline at 16 / 32 / 48 chunks, GPU tests
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
SODA_CUDA_TUNE_MIN_CELLS_LOG2=24 timeout 300 python tools/e2e_ab.py 16 24 32 > $O/r02w_e2e_old_threshold.jsonl 2> $O/r02w_e2e_old.err
echo "--- old threshold"; cat $O/r02w_e2e_old_threshold.jsonl; tail -2 $O/r02w_e2e_old.err
timeout 400 python tools/e2e_ab.py 16 24 32 48 64 -32 -48 96 32 16 > $O/r02w_e2e_new_threshold.jsonl 2> $O/r02w_e2e_new.err
echo "--- new threshold"; cat $O/r02w_e2e_new_threshold.jsonl; tail -2 $O/r02w_e2e_new.err
for T in 24 22; do
SODA_CUDA_TUNE_MIN_CELLS_LOG2=$T SODA_CUDA_PIPELINE_TRACE=1 timeout 300 python tools/e2e_ab.py 16 32 48 2> $O/r02w_trace_log2_$T.err > /dev/null
grep pipeline_trace $O/r02w_trace_log2_$T.err > $O/r02w_trace_log2_$T.jsonl
done
python - <<PY
import json
for t in (24, 22):
  lines = [json.loads(l)['pipeline_trace'] for l in open('$O/r02w_trace_log2_%d.jsonl' % t)]
  seen = set()
  for d in lines[::-1]:
    if d['chunks'] in seen: continue
    seen.add(d['chunks'])
    c = [e - b for b, e in zip(d['compute_begin_ms'], d['compute_end_ms'])]
    u = [e - b for b, e in zip(d['upload_begin_ms'], d['upload_end_ms'])]
    dn = [e - b for b, e in zip(d['download_begin_ms'], d['download_end_ms'])]
    mid = len(c) // 2
    print('log2', t, 'chunks', d['chunks'], 'total %.2f' % d['download_end_ms'][-1],
          'upload mid %.3f' % u[mid], 'compute mid %.3f first %.3f' % (c[mid], c[0]),
          'download mid %.3f' % dn[mid], 'first compute begins %.2f' % d['compute_begin_ms'][0],
          'last compute ends %.2f' % d['compute_end_ms'][-1])
PY
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02w_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02w_pytest_gpu.log
