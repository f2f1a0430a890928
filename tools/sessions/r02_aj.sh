#!/bin/bash
# round 2, session aj (2 GPUs): the final library -
# the pre-measured chunk windows - GPU tests of the slab path, the full bench
# line at N = 2 as the driver launches it
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -2
n=2
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2968$n bench.py --gpus $n --steps 10 --warmup 3 > $O/r02aj_bench_n$n.json 2> $O/r02aj_bench_n$n.err ) 2>&1 | grep real; echo "bench$n exit $?"
python - <<PY
import json
try:
  d=json.loads(open('$O/r02aj_bench_n$n.json').read().strip().splitlines()[0])
  print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d['parity'])
  e=d['e2e']; print('  e2e', e['value'], 'ms', e['ms_per_step'], 'peak', e['pcie_peak_gbs'], 'frac', e['frac'], 'parity', e['parity'])
  print('  c5', json.dumps(d['c5_strong'])[:700])
except Exception as e:
  print('ERR', e)
PY
tail -3 $O/r02aj_bench_n$n.err
