#!/bin/bash
# round 2, session aa (1 GPU): chunk lengths of the host pipeline that shrink
# towards the end (the call ends one chunk's passes and download after the last
# upload; a chunk may be at most ~8 % shorter than the one before it without
# the downloads queueing up).  SODA_CUDA_CHUNK_WEIGHTS layouts against 16 equal chunks
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
W() { python - "$@" <<PY
import sys
kind = sys.argv[1]
a = [float(x) for x in sys.argv[2:]]
if kind == 'decay':      # n chunks, each r times the one before
  n, r = int(a[0]), a[1]
  w = [r ** k for k in range(n)]
elif kind == 'flatdecay':  # m equal chunks, then n decaying by r
  m, n, r = int(a[0]), int(a[1]), a[2]
  w = [1.0] * m + [r ** (k + 1) for k in range(n)]
elif kind == 'growdecay':  # quarter, half, then n chunks decaying by r
  n, r = int(a[0]), a[1]
  w = [0.25, 0.5] + [r ** k for k in range(n)]
print(','.join('%.4f' % x for x in w))
PY
}
: > $O/r02aa_chunk_weights.jsonl
run() {
  local name="$1" weights="$2"
  SODA_CUDA_CHUNK_WEIGHTS="$weights" timeout 200 python tools/e2e_ab.py 16 2>> $O/r02aa.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); d['layout'] = '$name'; d['weights'] = '$weights'; del d['chunks']
print(json.dumps(d))" >> $O/r02aa_chunk_weights.jsonl
}
run "16 equal" ""
run "decay 24 x0.93" "$(W decay 24 0.93)"
run "decay 24 x0.90" "$(W decay 24 0.90)"
run "decay 32 x0.95" "$(W decay 32 0.95)"
run "decay 20 x0.92" "$(W decay 20 0.92)"
run "10 equal + 12 x0.88" "$(W flatdecay 10 12 0.88)"
run "12 equal + 10 x0.85" "$(W flatdecay 12 10 0.85)"
run "8 equal + 16 x0.92" "$(W flatdecay 8 16 0.92)"
run "quarter, half, 22 x0.93" "$(W growdecay 22 0.93)"
run "16 equal (again)" ""
python - <<PY
import json
for l in open('$O/r02aa_chunk_weights.jsonl'):
  d = json.loads(l); print('%-28s best %.2f mean %.2f ms' % (d['layout'], d['ms_best'], d['ms_mean']))
PY
tail -3 $O/r02aa.err
