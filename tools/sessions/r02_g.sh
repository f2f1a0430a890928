#!/bin/bash
# round 2, session g (4 GPUs): transport A/B at N = 2, bench at N = 4, copy
# peak of 4 GPUs at once, one process driving 2 and 4 devices
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
./tools/probe/pcie_probe 512 > $O/r02g_pcie_4gpu.jsonl 2>&1; cat $O/r02g_pcie_4gpu.jsonl
for t in nccl torch nccl torch; do
  SODA_BENCH_TRANSPORT=$t CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 30 --warmup 3 --headline-only 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t N=2', d['value'], d['ms_per_step'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r02g_bench_n4.json 2> $O/r02g_bench_n4.err; echo "bench4 exit $?"
python - <<PY
import json
try:
  d=json.loads(open('$O/r02g_bench_n4.json').read().strip().splitlines()[0])
  print('N=4 value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'])
  print('  e2e', json.dumps(d['e2e'])[:900])
  print('  c5', json.dumps(d['c5_strong'])[:900])
except Exception as e:
  print('ERR', e)
PY
tail -3 $O/r02g_bench_n4.err
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -3
timeout 300 python tools/multi_device_host.py > $O/r02g_multi_device_host.jsonl 2> $O/r02g_multi_device_host.err; cat $O/r02g_multi_device_host.jsonl; tail -3 $O/r02g_multi_device_host.err
