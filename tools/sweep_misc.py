#!/usr/bin/env python3
"""One-off sweep of launch-shape options for the 2-D programs other than the
bench (`build` compiles, `run` times on cuda:0; JSON lines like bench_configs)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    ('seidel2d', '16384,16384', 24, 3, {}), ('seidel2d', '16384,16384', 24, 4, {}),
    ('seidel2d', '16384,16384', 24, 6, {}), ('seidel2d', '16384,16384', 24, 4, {'warps': 4}),
    ('jacobi2d', '16384,16384', 120, 6, {'min_blocks': 5}),
    ('jacobi2d', '16384,16384', 120, 6, {'min_blocks': 6}),
    ('jacobi2d', '16384,16384', 120, 6, {'warps': 1}),
    ('jacobi2d', '16384,16384', 120, 6, {'stages': 2}),
    ('jacobi2d', '16384,16384', 120, 6, {}),
    ('denoise2d', '8192,8192', None, None, {}), ('denoise2d', '8192,8192', None, None, {'no_pipeline': True}),
    ('denoise2d', '8192,8192', None, None, {'warps': 2}), ('denoise2d', '8192,8192', None, None, {'warps': 8}),
    ('blur', '16000,16384', 2, 2, {}), ('blur', '16000,16384', 2, 2, {'warps': 2}),
    ('blur', '16000,16384', 2, 2, {'warps': 8}), ('blur', '16000,16384', 2, 2, {'no_pipeline': True}),
    ('blur', '16000,16384', 2, 2, {'cells': 16, 'stages': 3}),
    ('sobel2d', '16384,16384', None, None, {}), ('sobel2d', '16384,16384', None, None, {'warps': 2}),
    ('sobel2d', '16384,16384', None, None, {'cells': 16, 'stages': 3}),
    ('erosion', '16384,16384', None, None, {'warps': 2}), ('erosion', '16384,16384', None, None, {'no_pipeline': True}),
    ('xcorr', '16384,16384', None, None, {'warps': 2}), ('xcorr', '16384,16384', None, None, {'no_pipeline': True}),
]


def main():
  for name, extent, iterate, tb, options in CASES:
    cmd = [sys.executable, os.path.join(ROOT, 'tools', 'run_one.py'), name,
           extent, '--options', json.dumps(options)]
    if iterate:
      cmd += ['--iterate', str(iterate)]
    if tb:
      cmd += ['--tb', str(tb)]
    if sys.argv[1] == 'build':
      cmd.append('--build-only')
    out = subprocess.run(cmd, capture_output=True, text=True)
    print((out.stdout.strip() or out.stderr.strip()[-300:]).split('\n')[-1],
          flush=True)


if __name__ == '__main__':
  main()
