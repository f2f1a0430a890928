#!/bin/bash
# round 2, session ae (1 GPU): chunk counts of the host pipeline for the 3-D
# BASELINE programs (few passes, long reach: the 8x-reach rule left them
# unpipelined)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
{
timeout 300 python tools/e2e_any.py heat3d 512,512,512 --iterate 32 --chunks 1 2 3 4 6 8 0
timeout 300 python tools/e2e_any.py jacobi3d 512,512,512 --iterate 32 --chunks 1 4 0
timeout 300 python tools/e2e_any.py denoise3d 512,512,512 --chunks 1 16 0
timeout 300 python tools/e2e_any.py blur 2000,16384 --iterate 2 --chunks 1 16 0
timeout 300 python tools/e2e_any.py jacobi2d 16384,16384 --iterate 256 --chunks 1 4 8 0
SODA_CUDA_CHUNK_REACH=8 timeout 300 python tools/e2e_any.py jacobi2d 16384,16384 --iterate 256 --chunks 0
} > $O/r02ae_e2e_programs.jsonl 2> $O/r02ae.err
python - <<PY
import json
for l in open('$O/r02ae_e2e_programs.jsonl'):
  d = json.loads(l); print('%-10s %s iterate %d passes %d chunks %2d  best %.2f mean %.2f ms  %.1f GB/s' % (d['program'], d['extent'], d['iterate'], d['passes'], d['chunks'], d['ms_best'], d['ms_mean'], d['gbs_mean']))
PY
tail -3 $O/r02ae.err
