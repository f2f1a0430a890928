#!/bin/bash
# one 8-GPU box: the bench at N = 1, 2, 4, 8 (what the driver does at round end)
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 10 --warmup 3 > $O/scale_n$N.json 2> $O/scale_n$N.err
done
cat $O/scale_n1.json $O/scale_n2.json $O/scale_n4.json $O/scale_n8.json
