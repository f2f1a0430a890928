#!/bin/bash
# compute-sanitizer on small grids: memcheck (out-of-bounds, misaligned) and
# racecheck (shared-memory hazards of the one-barrier-per-plane protocol)
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name tool program extent extra...
  local name=$1 tool=$2; shift 2
  $CS --tool $tool --error-exitcode 9 python tools/run_one.py "$@" --reps 1 --warmup 0 > $O/sanitizer_${name}_${tool}.log 2>&1
  echo "$name $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitizer_${name}_${tool}.log | tail -1)"
}
run j3d_tb2 memcheck jacobi3d 150,37,29 --iterate 4 --tb 2
run j3d_tb2 racecheck jacobi3d 150,37,29 --iterate 4 --tb 2
run heat3d_tb3 racecheck heat3d 140,50,21 --iterate 3 --tb 3
run dn3d memcheck denoise3d 140,45,17
run dn3d racecheck denoise3d 140,45,17
run j2d_tb6 memcheck jacobi2d 1000,300 --iterate 12 --tb 6
run j2d_tb6 racecheck jacobi2d 1000,300 --iterate 12 --tb 6
run blur memcheck blur 2100,77 --iterate 2 --tb 2
run contrast memcheck contrast 600,90
$CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_stream_layout.py -m gpu -x -q > $O/sanitizer_layout_memcheck.log 2>&1
echo "layout memcheck rc=$? $(grep -E 'ERROR SUMMARY' $O/sanitizer_layout_memcheck.log | tail -1)"
