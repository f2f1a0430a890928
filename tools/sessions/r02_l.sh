#!/bin/bash
# round 2, session l (8 GPUs): the bench line at N = 8 (weak scaling, e2e
# through the slab host pipeline, C5 strong scaling, parity on every rank)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --steps 10 --warmup 3 --no-other > $O/r02l_bench_n8.json 2> $O/r02l_bench_n8.err; echo "bench8 exit $?"
python - <<PY
import json
try:
  d=json.loads(open('$O/r02l_bench_n8.json').read().strip().splitlines()[0])
  print('N=8 value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'])
  print('  e2e', json.dumps(d['e2e'])[:900])
  print('  c5', json.dumps(d['c5_strong'])[:900])
except Exception as e:
  print('ERR', e)
PY
tail -3 $O/r02l_bench_n8.err
./tools/probe/pcie_probe 512 > $O/r02l_pcie_8gpu.jsonl 2>&1; tail -6 $O/r02l_pcie_8gpu.jsonl
