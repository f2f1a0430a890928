#!/bin/bash
# 8 GPUs: N-GPU == 1-GPU bit for bit (tools/multi_gpu_check.py)
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 tools/multi_gpu_check.py > $O/multi_gpu_check_n8.jsonl 2> $O/multi_gpu_check_n8.err
cat $O/multi_gpu_check_n8.jsonl; tail -3 $O/multi_gpu_check_n8.err
