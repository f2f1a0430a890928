#!/bin/bash
# round 2, session f (2 GPUs): segment tuning variance on the headline kernel,
# halo transport comparison at N = 2
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for i in 1 2 3; do
  SODA_CUDA_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --headline-only 2> $O/r02f_h$i.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('auto', d['value'], d['ms_per_step'])"
  grep "segment" $O/r02f_h$i.err | head -4
done
for seg in 300 400 497 600 800; do
  SODA_CUDA_AUTOTUNE=0 SODA_CUDA_SEGMENT=$seg timeout 300 python bench.py --steps 20 --warmup 3 --headline-only 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('segment $seg', d['value'], d['ms_per_step'])"
done
for t in nccl torch; do
  SODA_BENCH_TRANSPORT=$t NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 --no-other --no-cpu-baseline > $O/r02f_bench_n2_$t.json 2> $O/r02f_bench_n2_$t.err; echo "bench2 $t exit $?"
  python - <<PY
import json
try:
  d=json.loads(open('$O/r02f_bench_n2_$t.json').read().strip().splitlines()[0])
  print('$t', 'value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity']['bit_exact'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('parity',{}).get('bit_exact'))
  print('  c5', json.dumps(d['c5_strong'])[:700])
except Exception as e:
  print('$t', 'ERR', e)
PY
  tail -3 $O/r02f_bench_n2_$t.err
done
