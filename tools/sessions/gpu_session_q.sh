#!/bin/bash
# 2 GPUs: N-GPU == 1-GPU bit for bit for 3-D, multi-input, uint16 and grouped-exchange cases
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/multi_gpu_check.py > $O/multi_gpu_check_n2.jsonl 2> $O/multi_gpu_check_n2.err
cat $O/multi_gpu_check_n2.jsonl; tail -3 $O/multi_gpu_check_n2.err
