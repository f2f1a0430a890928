#!/bin/bash
# round 2, session b (2 GPUs): slab runtime behind the ABI - GPU tests, bitwise
# N-GPU == 1-GPU check, bench at N = 1 and 2, copy peak of two GPUs at once
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py -x -q > $O/r02b_pytest.log 2>&1; echo "pytest exit $?"; tail -15 $O/r02b_pytest.log
./tools/probe/pcie_probe 512 > $O/r02b_pcie_2gpu.jsonl 2>&1; cat $O/r02b_pcie_2gpu.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/multi_gpu_check.py > $O/r02b_multi_gpu_check_n2.jsonl 2> $O/r02b_multi_gpu_check_n2.err; echo "check exit $?"
cat $O/r02b_multi_gpu_check_n2.jsonl | cut -c1-300; tail -5 $O/r02b_multi_gpu_check_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02b_bench_n2.json 2> $O/r02b_bench_n2.err; echo "bench2 exit $?"
tail -c 2500 $O/r02b_bench_n2.json; tail -8 $O/r02b_bench_n2.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-other > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; echo "bench1 exit $?"
tail -c 1500 $O/r02b_bench_n1.json; tail -5 $O/r02b_bench_n1.err
