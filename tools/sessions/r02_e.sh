#!/bin/bash
# round 2, session e (1 GPU): GPU tests, bench line, launch list and ncu
# captures of the shipped kernels, blur CTA shapes on the wide grid
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02e_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $O/r02e_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02e_bench.json 2> $O/r02e_bench.err; echo "bench exit $?"
tail -c 600 $O/r02e_bench.json; tail -5 $O/r02e_bench.err
R="timeout 300 python tools/run_one.py"
{
for w in 1 2 4; do
  $R blur 16000,16384 --iterate 2 --tb 2 --options "{\"warps\":$w}"
done
$R blur 2000,16384 --iterate 2 --tb 2
} > $O/r02e_blur.jsonl 2> $O/r02e_blur.err
cut -c1-220 $O/r02e_blur.jsonl
timeout 600 ncu -k regex:soda --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02e_bench_launches.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/r02e_ncu_list.log 2>&1; echo "ncu list exit $?"
SODA_CUDA_AUTOTUNE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:soda_stream2d --launch-skip 13 --launch-count 1 -o $O/r02e_prof_j2d_tb6 python bench.py --steps 2 --warmup 1 --headline-only > $O/r02e_ncu_full.log 2>&1; echo "ncu full exit $?"
for p in jacobi3d heat3d; do
  SODA_CUDA_AUTOTUNE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:soda_stream3d -s 1 -c 1 -o $O/r02e_prof_${p}_tb2 python tools/run_one.py $p 512,512,512 --iterate 32 --tb 2 --reps 1 --warmup 1 > $O/r02e_ncu_$p.log 2>&1; echo "ncu $p exit $?"
done
SODA_CUDA_AUTOTUNE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:soda_stream2d -s 1 -c 1 -o $O/r02e_prof_blur_c1 python tools/run_one.py blur 2000,16384 --iterate 2 --tb 2 --reps 1 --warmup 1 > $O/r02e_ncu_blur.log 2>&1; echo "ncu blur exit $?"
