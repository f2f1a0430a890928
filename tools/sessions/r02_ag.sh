#!/bin/bash
# round 2, session ag (1 GPU): dense rows / planes as one linear copy instead of
# a pitched one (uploads of whole rows; 3-D programs whose outputs have no
# dimension-0 / 1 border never qualify on the way back) - e2e on and off,
# alternating; then the final validation of session af again
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r02ag_linear_copies.jsonl
run() {
  local tag="$1"; shift
  SODA_CUDA_LINEAR_COPIES=$tag timeout 300 python "$@" 2>> $O/r02ag.err | python -c "
import sys, json
for l in sys.stdin:
  d = json.loads(l); d['linear_copies'] = $tag; print(json.dumps(d))" >> $O/r02ag_linear_copies.jsonl
}
for rep in 1 2 3; do
run 0 tools/e2e_ab.py 0
run 1 tools/e2e_ab.py 0
done
run 0 tools/e2e_any.py heat3d 512,512,512 --iterate 32 --chunks 0
run 1 tools/e2e_any.py heat3d 512,512,512 --iterate 32 --chunks 0
run 0 tools/e2e_any.py denoise3d 512,512,512 --chunks 0
run 1 tools/e2e_any.py denoise3d 512,512,512 --chunks 0
python - <<PY
import json
for l in open('$O/r02ag_linear_copies.jsonl'):
  d = json.loads(l); print(d.get('program', 'jacobi2d x64'), 'linear' if d['linear_copies'] else 'pitched', 'best %.2f mean %.2f ms' % (d['ms_best'], d['ms_mean']))
PY
tail -3 $O/r02ag.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02ag_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02ag_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 700 python bench.py --steps 20 --warmup 3 > $O/r02ag_bench.json 2> $O/r02ag_bench.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02ag_bench_reference.json 2> $O/r02ag_bench_reference.err; echo "reference exit $?"; cut -c1-400 $O/r02ag_bench_reference.json
python - <<PY
import json
d=json.loads(open('$O/r02ag_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'lib', d['roofline']['library'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['frac'])
print('dataflow', json.dumps(d.get('cpu_baseline_dataflow'))[:300])
for c in d['other_configs']:
  print(c.get('config'), c.get('value'), c.get('roofline',{}).get('frac'), c.get('parity',{}).get('bit_exact'), c.get('error'))
PY
tail -3 $O/r02ag_bench.err
timeout 600 ncu -k regex:soda --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02ag_bench_launches.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/r02ag_ncu_list.log 2>&1; echo "ncu list exit $?"
SODA_CUDA_AUTOTUNE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:soda_stream2d --launch-skip 13 --launch-count 1 -o $O/r02ag_prof_j2d_tb6 python bench.py --steps 2 --warmup 1 --headline-only > $O/r02ag_ncu_full.log 2>&1; echo "ncu full exit $?"
