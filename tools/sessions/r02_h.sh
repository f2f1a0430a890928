#!/bin/bash
# round 2, session h (1 GPU): GPU tests after the host-pipeline change and the
# wide-integer programs, chunk layouts of the e2e path, full bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/e2e_ab.py > $O/r02h_e2e_ab.jsonl 2> $O/r02h_e2e_ab.err; cat $O/r02h_e2e_ab.jsonl; tail -3 $O/r02h_e2e_ab.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02h_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $O/r02h_pytest_gpu.log
timeout 700 python bench.py --steps 20 --warmup 3 > $O/r02h_bench.json 2> $O/r02h_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('$O/r02h_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print('e2e', json.dumps(d['e2e'])[:700])
print('dataflow', json.dumps(d.get('cpu_baseline_dataflow'))[:500])
for c in d['other_configs']:
  print(c.get('config'), c.get('value'), c.get('roofline',{}).get('frac'), c.get('error'))
PY
tail -3 $O/r02h_bench.err
timeout 900 python tools/random_sweep.py run 40 100 > $O/r02h_random_sweep.jsonl 2> $O/r02h_random_sweep.err; echo "sweep exit $?"; tail -1 $O/r02h_random_sweep.jsonl; grep -v '"status": "ok"' $O/r02h_random_sweep.jsonl | cut -c1-200 | head -5
timeout 900 python tools/random_sweep.py run 0 40 --hard > $O/r02h_random_sweep_hard.jsonl 2> $O/r02h_random_sweep_hard.err; echo "hard sweep exit $?"; tail -1 $O/r02h_random_sweep_hard.jsonl; grep -v '"status": "ok"' $O/r02h_random_sweep_hard.jsonl | cut -c1-200 | head -5
