#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python tools/tune2d.py run > $O/tune2d_v4.log 2>&1
SODA_TUNE_SET=tb2 python tools/tune3d.py run > $O/tune3d_v2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:soda_stream3d -s 1 -c 1 -o $O/prof_j3d_tb2_v2 -f python tools/run_one.py jacobi3d 512,512,512 --iterate 4 --tb 2 --options '{"rows": 32, "cy": 4, "min_blocks": 2}' --reps 1 --warmup 1 > $O/ncu_j3d_tb2_v2.log 2>&1
$NCU -k regex:soda_stream3d -s 2 -c 1 -o $O/prof_j3d_tb1_v2 -f python tools/run_one.py jacobi3d 512,512,512 --iterate 4 --tb 1 --options '{"rows": 8, "cy": 2}' --reps 1 --warmup 1 > $O/ncu_j3d_tb1_v2.log 2>&1
tail -3 $O/tune2d_v4.log $O/tune3d_v2.log
