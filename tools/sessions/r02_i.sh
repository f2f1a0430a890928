#!/bin/bash
# round 2, session i (4 GPUs): slab / multi-device tests and bench at N = 2, 4
# after the host-pipeline change (upload pieces end where chunk windows end)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_fixed_point.py tests/test_params.py -m gpu -x -q 2>&1 | tail -3
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2966$n bench.py --gpus $n --steps 10 --warmup 3 --no-other > $O/r02i_bench_n$n.json 2> $O/r02i_bench_n$n.err; echo "bench$n exit $?"
  python - <<PY
import json
try:
  d=json.loads(open('$O/r02i_bench_n$n.json').read().strip().splitlines()[0])
  print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'])
  print('  e2e', json.dumps(d['e2e'])[:900])
  print('  c5', json.dumps(d['c5_strong'])[:900])
except Exception as e:
  print('ERR', e)
PY
  tail -3 $O/r02i_bench_n$n.err
done
timeout 300 python tools/multi_device_host.py > $O/r02i_multi_device_host.jsonl 2> $O/r02i_multi_device_host.err; cat $O/r02i_multi_device_host.jsonl; tail -3 $O/r02i_multi_device_host.err
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 tools/multi_gpu_check.py > $O/r02i_multi_gpu_check_n2.jsonl 2> $O/r02i_multi_gpu_check_n2.err; echo "check exit $?"; cut -c1-160 $O/r02i_multi_gpu_check_n2.jsonl | tail -9
