#!/bin/bash
# round 2, session d (1 GPU): 3-D experiments (packed pairs, tile shapes) and
# blur launch shapes for the narrow C1 grid
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
R="timeout 300 python tools/run_one.py"
{
for p in jacobi3d heat3d; do
  for tb in 2 3 4; do
    $R $p 512,512,512 --iterate $((tb*4)) --tb $tb
    $R $p 512,512,512 --iterate $((tb*4)) --tb $tb --options '{"pack": true}'
  done
  $R $p 512,512,512 --iterate 12 --tb 3 --options '{"rows": 32, "cy": 2}'
  $R $p 512,512,512 --iterate 12 --tb 3 --options '{"rows": 32, "cy": 4}'
  $R $p 512,512,512 --iterate 12 --tb 3 --options '{"rows": 32, "cy": 4, "pack": true}'
  $R $p 512,512,512 --iterate 16 --tb 4 --options '{"rows": 32, "cy": 4, "pack": true}'
  $R $p 512,512,512 --iterate 16 --tb 4 --options '{"rows": 40, "cy": 4, "pack": true}'
  $R $p 512,512,512 --iterate 8 --tb 2 --options '{"rows": 32, "cy": 4}'
  $R $p 512,512,512 --iterate 8 --tb 2 --options '{"rows": 32, "cy": 4, "pack": true}'
done
for w in 1 2 3 4; do
  $R blur 2000,16384 --iterate 2 --tb 2 --options "{\"warps\": $w}"
  $R blur 2000,16384 --iterate 2 --tb 1 --options "{\"warps\": $w}"
done
$R denoise3d 512,512,512
$R denoise3d 512,512,512 --options '{"rows": 12}'
$R denoise3d 512,512,512 --options '{"rows": 16, "cy": 2}'
} > $O/r02d_experiments.jsonl 2> $O/r02d_experiments.err
cut -c1-260 $O/r02d_experiments.jsonl; tail -5 $O/r02d_experiments.err
