#!/bin/bash
# round 2, session s (1 GPU): equal-chunk counts of the e2e pipeline after the
# upload-piece change
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python tools/e2e_ab.py 16 20 24 32 12 16 > $O/r02s_e2e_chunks.jsonl 2> $O/r02s_e2e_chunks.err; cat $O/r02s_e2e_chunks.jsonl; tail -3 $O/r02s_e2e_chunks.err
