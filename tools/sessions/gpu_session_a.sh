#!/bin/bash
# GPU session: 2-D tuning sweep, then ncu captures of the 3-D kernels.
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu3.log 2>&1; tail -3 $O/pytest_gpu3.log
python tools/tune2d.py run > $O/tune2d_v3.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
python tools/run_one.py jacobi3d 512,512,512 --iterate 4 --tb 1 --options '{"rows": 8}' > $O/j3d_tb1.json 2>&1 && \
$NCU -k regex:soda_stream3d -s 2 -c 1 -o $O/prof_j3d_tb1 -f python tools/run_one.py jacobi3d 512,512,512 --iterate 4 --tb 1 --options '{"rows": 8}' --reps 1 --warmup 1 > $O/ncu_j3d_tb1.log 2>&1
python tools/run_one.py jacobi3d 512,512,512 --iterate 4 --tb 2 --options '{"rows": 32}' > $O/j3d_tb2.json 2>&1 && \
$NCU -k regex:soda_stream3d -s 1 -c 1 -o $O/prof_j3d_tb2 -f python tools/run_one.py jacobi3d 512,512,512 --iterate 4 --tb 2 --options '{"rows": 32}' --reps 1 --warmup 1 > $O/ncu_j3d_tb2.log 2>&1
python tools/run_one.py denoise3d 512,512,512 --options '{"rows": 16}' > $O/dn3d.json 2>&1 && \
$NCU -k regex:soda_stream3d -s 1 -c 1 -o $O/prof_dn3d -f python tools/run_one.py denoise3d 512,512,512 --options '{"rows": 16}' --reps 1 --warmup 1 > $O/ncu_dn3d.log 2>&1
python tools/run_one.py blur 16000,16384 --iterate 2 --tb 2 > $O/blur.json 2>&1 && \
$NCU -k regex:soda_stream2d -s 1 -c 1 -o $O/prof_blur -f python tools/run_one.py blur 16000,16384 --iterate 2 --tb 2 --reps 1 --warmup 1 > $O/ncu_blur.log 2>&1
cat $O/j3d_tb1.json $O/j3d_tb2.json $O/dn3d.json $O/blur.json
ls -la $O/*.ncu-rep
