"""TEST INFRASTRUCTURE: compiles a generated SODA CUDA program for the CPU
emulation in soda_emu.h (g++ -DSODA_EMU), so the real kernel templates, plan
tables and host runtime can be exercised without a GPU.  Never used by the
product."""
import hashlib
import os
import subprocess

from soda_b200.codegen.cuda import build, emit

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, '_build')


def build_emu_library(stencil, time_block=None, options=None,
                      sanitize=False) -> str:
  source = emit.emit_program(stencil, time_block=time_block, options=options)
  with open(os.path.join(HERE, 'soda_emu.h'), 'rb') as fp:
    emu_header = fp.read()
  digest = hashlib.sha1(source.encode() + emu_header +
                        build._headers_digest().encode() +
                        str(sanitize).encode()).hexdigest()[:12]
  os.makedirs(BUILD_DIR, exist_ok=True)
  base = os.path.join(BUILD_DIR, 'emu_%s_%s' % (stencil.app_name, digest))
  lib = base + '.so'
  if os.path.exists(lib):
    return lib
  with open(base + '.cpp', 'w') as fp:
    fp.write(source)
  cmd = [
      'g++', '-std=c++20', '-O1', '-g', '-shared', '-fPIC', '-pthread',
      '-DSODA_EMU', '-ffp-contract=off', '-fvisibility=hidden', '-fno-gnu-unique', '-Wno-unknown-pragmas', '-I', HERE,
      '-I', build.CSRC_DIR, '-I', build.INCLUDE_DIR, base + '.cpp', '-o',
      lib + '.tmp'
  ]
  if sanitize:
    cmd[1:1] = ['-fsanitize=address,undefined', '-fno-omit-frame-pointer']
  result = subprocess.run(cmd, capture_output=True, text=True)
  if result.returncode != 0:
    raise RuntimeError('emulation build failed:\n%s' % result.stderr[-6000:])
  os.replace(lib + '.tmp', lib)
  return lib
