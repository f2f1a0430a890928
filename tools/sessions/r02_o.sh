#!/bin/bash
# round 2, session o (1 GPU): tile heights that fit 512 rows better
# (valid rows 12 of 16 -> 43 tiles = 688 row slots; 16 of 20 -> 640; 20 of 24 -> 624)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
R="timeout 300 python tools/run_one.py"
{
for p in heat3d jacobi3d; do
  $R $p 512,512,512 --iterate 32 --tb 2
  for rows in 20 24 28; do
    $R $p 512,512,512 --iterate 32 --tb 2 --options "{\"rows\": $rows, \"cy\": 4}"
  done
  $R $p 512,512,512 --iterate 32 --tb 2 --options '{"rows": 20, "cy": 2}'
  $R $p 512,512,512 --iterate 32 --tb 2 --options '{"rows": 18, "cy": 2}'
  $R $p 512,512,512 --iterate 32 --tb 2 --options '{"rows": 20, "cy": 5}'
done
} > $O/r02o_tile_rows.jsonl 2> $O/r02o_tile_rows.err
cut -c1-260 $O/r02o_tile_rows.jsonl; tail -3 $O/r02o_tile_rows.err
