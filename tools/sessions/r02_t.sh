#!/bin/bash
# round 2, session t (2 GPUs): slab run == 1-GPU run, bit for bit, on the seeded
# random programs (plain 0-39, hard 0-39), 1-3 passes per halo exchange
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29673 tools/multi_gpu_check.py random 0 40 > $O/r02t_multi_gpu_random_n2.jsonl 2> $O/r02t_multi_gpu_random_n2.err; echo "plain exit $?"; tail -1 $O/r02t_multi_gpu_random_n2.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29674 tools/multi_gpu_check.py random 0 40 hard > $O/r02t_multi_gpu_random_hard_n2.jsonl 2> $O/r02t_multi_gpu_random_hard_n2.err; echo "hard exit $?"; tail -1 $O/r02t_multi_gpu_random_hard_n2.jsonl
tail -3 $O/r02t_multi_gpu_random_hard_n2.err
