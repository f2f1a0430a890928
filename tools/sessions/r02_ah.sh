#!/bin/bash
# round 2, session ah (1 GPU): final validation at the last change of the runtime (chunk windows measured once per process and layout) - GPU tests, smoke, bench line of
# both arms, ncu launch list and full captures of the shipped kernels
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02ah_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02ah_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 700 python bench.py --steps 20 --warmup 3 > $O/r02ah_bench.json 2> $O/r02ah_bench.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02ah_bench_reference.json 2> $O/r02ah_bench_reference.err; echo "reference exit $?"; cut -c1-400 $O/r02ah_bench_reference.json
python - <<PY
import json
d=json.loads(open('$O/r02ah_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'lib', d['roofline']['library'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['frac'])
print('dataflow', json.dumps(d.get('cpu_baseline_dataflow'))[:300])
for c in d['other_configs']:
  print(c.get('config'), c.get('value'), c.get('roofline',{}).get('frac'), c.get('parity',{}).get('bit_exact'), c.get('error'))
PY
tail -3 $O/r02ah_bench.err
timeout 600 ncu -k regex:soda --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02ah_bench_launches.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/r02ah_ncu_list.log 2>&1; echo "ncu list exit $?"
SODA_CUDA_AUTOTUNE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:soda_stream2d --launch-skip 13 --launch-count 1 -o $O/r02ah_prof_j2d_tb6 python bench.py --steps 2 --warmup 1 --headline-only > $O/r02ah_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 300 python tools/e2e_any.py jacobi2d 16384,16384 --iterate 64 --chunks 0 --one-shot > $O/r02ah_one_shot.jsonl 2> $O/r02ah_one_shot.err; cat $O/r02ah_one_shot.jsonl | cut -c1-220
