#!/bin/bash
# round 2, session ac (1 GPU): the host pipeline measures its chunk windows
# before it queues the first upload - spread of the e2e time over processes;
# GPU tests; one bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r02ac_e2e.jsonl
run() {
  local name="$1" decay="$2" chunks="$3"
  SODA_CUDA_CHUNK_DECAY="$decay" timeout 200 python tools/e2e_ab.py $chunks 2>> $O/r02ac.err | python -c "
import sys, json
for l in sys.stdin:
  d = json.loads(l); d['layout'] = '$name'; print(json.dumps(d))" >> $O/r02ac_e2e.jsonl
}
for rep in 1 2 3 4; do
run "auto: 20 x0.92" 0.92 0
run "16 equal" 1 16
done
run "20 equal" 1 20
python - <<PY
import json
for l in open('$O/r02ac_e2e.jsonl'):
  d = json.loads(l); print('%-18s best %.2f mean %.2f ms' % (d['layout'], d['ms_best'], d['ms_mean']))
PY
tail -3 $O/r02ac.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02ac_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02ac_pytest_gpu.log
( time timeout 700 python bench.py --steps 20 --warmup 3 > $O/r02ac_bench.json 2> $O/r02ac_bench.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('$O/r02ac_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'lib', d['roofline']['library'], d['clocks'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['frac'], d['e2e']['parity'])
for c in d['other_configs']:
  print(c.get('config'), c.get('value'), c.get('roofline',{}).get('frac'), c.get('parity',{}).get('bit_exact'), c.get('error'))
print('c5', d['c5_strong'].get('value'), d['c5_strong'].get('parity'))
PY
tail -3 $O/r02ac_bench.err
