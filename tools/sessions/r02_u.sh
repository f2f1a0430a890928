#!/bin/bash
# round 2, session u (4 GPUs): end-to-end at N = 2 and 4 with the chunks that
# read ghost slices computed in their natural place (NCCL transport)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -2
for n in ${SODA_SESSION_NS:-2 4}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2968$n bench.py --gpus $n --steps 10 --warmup 3 --no-other --no-c5 > $O/r02u_bench_n$n.json 2> $O/r02u_bench_n$n.err; echo "bench$n exit $?"
  python - <<PY
import json
try:
  d=json.loads(open('$O/r02u_bench_n$n.json').read().strip().splitlines()[0])
  print('N=$n value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity']['bit_exact'])
  e=d['e2e']; print('  e2e', e['value'], 'ms', e['ms_per_step'], 'peak', e['pcie_peak_gbs'], 'frac', e['frac'], 'parity', e['parity']['bit_exact'])
except Exception as e:
  print('ERR', e)
PY
  tail -2 $O/r02u_bench_n$n.err
done
