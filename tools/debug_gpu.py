import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from soda_b200 import sodac
from soda_b200.codegen import cuda as cb
from soda_b200.codegen.cuda import launcher
from oracle import golden
from tests import common

def case(name, extent, tb=None, options=None, **ov):
    st = common.stencil(name, **ov)
    prog = cb.compile_stencil(st, time_block=tb, options=options)
    ins = common.make_inputs(st, extent, seed=1)
    out = prog.run_host(ins)
    want = common.oracle_outputs(st, ins)
    common.assert_matches_oracle(st, extent, out, want)
    print('OK', name, extent, tb, options, ov, flush=True)

which = sys.argv[1]
if which == 'j1':
    case('jacobi2d', (500, 300), tb=1)
elif which == 'j1big':
    case('jacobi2d', (16384, 1024), tb=1, iterate=2)
elif which == 'blur':
    case('blur', (2000, 4))
elif which == 'blur64':
    case('blur', (2000, 64))
elif which == 'sobel':
    case('sobel2d', (32, 4))
elif which == 'blur8':
    case('blur', (2000, 64), options={'cells': 8})
elif which == 'j2cells':
    case('jacobi2d', (500, 300), options={'cells': 2})
elif which == 'dev':
    import torch
    st = common.stencil('jacobi2d', iterate=16)
    prog = cb.compile_stencil(st, time_block=1, options={'warps': 4, 'chunk': 6, 'stages': 3})
    d_in = torch.rand((16384, 16384), dtype=torch.float32, device='cuda')
    d_out = torch.zeros_like(d_in)
    plan = prog.create_plan((16384, 16384), launcher.make_opts(stream=torch.cuda.current_stream().cuda_stream))
    plan.run_device([d_in.data_ptr()], [(16384, 0)], [d_out.data_ptr()], [(16384, 0)])
    torch.cuda.synchronize()
    print('OK dev', flush=True)
if which == 'crseidel':
    for opts in ({'no_pack': True}, None):
        try:
            case('seidel2d', (500, 200), options=opts, computation_reuse='yes')
        except AssertionError as e:
            print('FAIL', opts, str(e)[:200])
    # same arithmetic shape without CR: 3-term sum times .1111111f
    import tempfile
    from soda_b200 import sodac as _s
    src = common.source('jacobi2d').replace('0.2f', '.1111111f')
    st = _s.compile_source(src)
    for opts in ({'no_pack': True}, None):
        prog = cb.compile_stencil(st, options=opts)
        ins = common.make_inputs(st, (500, 200), seed=1)
        out = prog.run_host(ins)
        want = common.oracle_outputs(st, ins)
        try:
            common.assert_matches_oracle(st, (500, 200), out, want)
            print('OK jacobi .1111111f', opts)
        except AssertionError as e:
            print('FAIL jacobi .1111111f', opts, str(e)[:200])
