#!/bin/bash
# BASELINE config C5 on one 8-GPU box: 65536^2 x 256, strong scaling at 8 GPUs,
# bit-compared with the 1-GPU run of the whole grid and with the oracle
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/c5_scaling.py > $O/c5_n8.json 2> $O/c5_n8.err
cat $O/c5_n8.json; tail -3 $O/c5_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 tools/c5_scaling.py > $O/c5_n4.json 2> $O/c5_n4.err
cat $O/c5_n4.json
