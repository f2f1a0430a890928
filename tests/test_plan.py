"""Planner tables: lags, register-window depths, halos, pass schedule."""
import pytest

from soda_b200 import sodac, util
from soda_b200.codegen.cuda import emit, plan
from tests import common


def test_jacobi2d_time_block_4():
  p = plan.make_pass_plan(common.stencil('jacobi2d', iterate=64), time_block=4)
  assert [n.name for n in p.nodes][:2] == ['t1', 't1_iter1']
  # pipelined schedule (2-D default): a level reads what the level below had
  # finished before the current step, so it trails it by 1 (row +1) + 1 step
  assert p.skew == 1
  assert [n.lag for n in p.nodes] == [0, 2, 4, 6, 8]
  # every level but the last is read at rows -1, 0, +1 by the next level
  assert [n.ring for n in p.nodes] == [3, 3, 3, 3, 1]
  assert p.nodes[-1].out == 0 and all(n.out < 0 for n in p.nodes[:-1])
  assert (p.cells, p.strip) == (4, 128)
  assert p.halo_lo == (4,) and p.halo_hi == (4,) and p.valid == (120,)
  assert (p.lo_s, p.max_lag) == (-4, 8)
  # producers-first schedule: the closed form of the reference's produce
  # offsets (src/soda/core.py:371-446)
  q = plan.make_pass_plan(common.stencil('jacobi2d', iterate=64), time_block=4,
                          pipelined=False)
  assert q.skew == 0 and [n.lag for n in q.nodes] == [0, 1, 2, 3, 4]
  assert [n.ring for n in q.nodes] == [3, 3, 3, 3, 1]
  assert (q.lo_s, q.max_lag) == (-4, 4)


def test_one_sided_window_blur():
  p = plan.make_pass_plan(common.stencil('blur'), pipelined=False)
  # uint16: 8 cells per lane (16-byte vectors, 16-byte aligned TMA boxes)
  assert (p.cells, p.strip, p.align0) == (8, 256, 8)
  assert p.halo_lo == (0,) and p.halo_hi == (2,)
  assert p.valid == (248,)
  by_name = {n.name: n for n in p.nodes}
  assert by_name['blur_x'].lag == 2 and by_name['blur_y'].lag == 2
  assert by_name['input'].ring == 3 and by_name['blur_x'].ring == 1


def test_off_centre_store_xcorr():
  """tmp1(0, 9) = sum input(0, 0..18): offsets are taken relative to the store
  index (src/soda/codegen/frt/host.py:587-592)."""
  p = plan.make_pass_plan(common.stencil('xcorr'), pipelined=False)
  by_name = {n.name: n for n in p.nodes}
  assert by_name['tmp1'].lag == 9 and by_name['input'].ring == 19
  assert by_name['tmp2'].halo_lo == (9,) and by_name['tmp2'].halo_hi == (9,)
  assert by_name['tmp3'].win_lo == (-9, -9) and \
      by_name['tmp3'].win_hi == (9, 9)


def test_3d_shared_memory_planes():
  p = plan.make_pass_plan(common.stencil('jacobi3d', iterate=32), time_block=2,
                          rows=8)
  t1, mid, out = p.nodes
  # dimension-1 neighbours of the input come from the TMA ring (no skew); those
  # of the fused intermediate from an exported plane, read one step late
  assert t1.smem_depth == 2 and mid.smem_depth == 2
  assert (t1.lag, mid.lag, out.lag) == (0, 1, 2)
  assert p.valid == (120, 4) and p.halo_lo == (4, 2)


def test_multi_input_dag_denoise3d():
  p = plan.make_pass_plan(common.stencil('denoise3d'), rows=8)
  by_name = {n.name: n for n in p.nodes}
  assert by_name['f'].kind == 'input' and by_name['u'].kind == 'input'
  out = by_name['output']
  assert [p.nodes[i].name for i in out.prods] == ['u', 'g', 'f', 'r1']
  assert out.win_lo == (-2, -2, -2) and out.win_hi == (2, 2, 2)
  assert by_name['g'].smem_depth >= 2  # g(0, +-1, 0) crosses warps


def test_pass_schedule_and_time_block_choice():
  assert plan.pass_schedule(64, 4) == [4] * 16
  assert plan.pass_schedule(10, 4) == [4, 4, 2]
  assert plan.pass_schedule(1, 1) == [1]
  assert plan.choose_time_block(common.stencil('denoise2d')) == 1
  assert plan.choose_time_block(common.stencil('jacobi2d')) == 2
  # the throughput model's choices (model.py; validated on B200 in
  # tests/test_model.py): the smallest time block within 12 % of the best -
  # 6 fused iterations in 2-D (8 is 9 % faster but FMA-bound), 2 in 3-D
  assert plan.choose_time_block(common.stencil('jacobi2d', iterate=64)) == 6
  assert plan.choose_time_block(common.stencil('heat3d', iterate=32)) == 2
  assert plan.choose_time_block(common.stencil('jacobi3d', iterate=32)) == 2
  assert plan.choose_time_block(common.stencil('heat3d', iterate=32), 3) == 3


def test_window_too_wide_for_a_strip():
  with pytest.raises(util.SemanticError):
    plan.make_pass_plan(common.stencil('jacobi2d', iterate=128),
                        time_block=70)


def test_generated_source_contains_no_kernels():
  """Only functors and tables are generated; kernels are the templates."""
  text = emit.emit_program(common.stencil('jacobi2d', iterate=64), 4)
  assert '__global__' not in text and '<<<' not in text
  assert 'soda::NodeDesc kNodes' in text and 'struct Stage<0>' in text
  assert 'a.template ld<0, -1, 0, 0>()' in text
  assert 'extern "C" SODA_CUDA_API int soda_cuda_jacobi2d(' in text


def test_packed_fp32_eligibility():
  """Packed FADD2/FFMA2 evaluation: fp32 programs that only add, subtract
  and multiply."""
  want = {'jacobi2d': True, 'jacobi3d': True, 'heat3d': True,
          'seidel2d': False,  # nine loads, six at dimension-0 offsets
          'blur': False, 'sobel2d': False, 'xcorr': False, 'erosion': False,
          'denoise2d': False, 'denoise3d': False,   # sqrt, division
          'contrast': False}
  for name, flag in want.items():
    assert plan.packable(common.stencil(name)) == flag, name
  p = plan.make_pass_plan(common.stencil('jacobi2d'), time_block=2)
  assert p.pack == 2
  assert plan.make_pass_plan(common.stencil('jacobi2d'), time_block=2,
                             pack=False).pack == 1
  text = emit.emit_program(common.stencil('jacobi2d'))
  assert 'kPack = 2' in text and 'soda::cast_to<float>' in text


def test_stages_run_as_late_as_their_consumers_allow():
  """ALAP: a stage nobody reads yet would only sit in a register window."""
  p = plan.make_tuned_pass_plan(common.stencil('denoise3d'), 1)
  by_name = {n.name: n for n in p.nodes}
  for name in ('diff_u', 'diff_d', 'diff_l', 'diff_r', 'diff_i', 'diff_o'):
    assert by_name[name].lag == by_name['g'].lag and by_name[name].ring == 1
  assert by_name['r1'].lag == by_name['output'].lag and by_name['r1'].ring == 1
  # stored nodes keep their as-soon-as-possible lag: the pass latency is the
  # reach of the window (2 planes)
  assert by_name['output'].lag == 2 and p.max_lag == 2
  # every consumer still runs behind its producers
  for node in p.nodes:
    for prod, deltas in zip(node.prods, node.deltas):
      for delta in deltas:
        assert node.lag - p.nodes[prod].lag - delta[-1] >= 0


def test_launch_shape_follows_the_dag():
  """make_tuned_pass_plan: wide lanes / patches while the register windows
  stay small (defaults measured on B200)."""
  j2 = plan.make_tuned_pass_plan(common.stencil('jacobi2d', iterate=64), 6)
  assert (j2.cells, j2.strip, j2.valid, j2.pack, j2.skew) == (8, 256, (240,), 2, 1)
  # eight fused iterations: 200 window registers at 8 cells, the budget of
  # packed programs (252 registers in all, no spill; measured faster than 4
  # cells); nine would need 224: back to 4-cell lanes
  assert plan.make_tuned_pass_plan(common.stencil('jacobi2d', iterate=64),
                                   8).cells == 8
  assert plan.make_tuned_pass_plan(common.stencil('jacobi2d', iterate=72),
                                   9).cells == 4
  # scalar arithmetic keeps the tighter budget (seidel2d is not packed)
  assert plan.make_tuned_pass_plan(common.stencil('seidel2d', iterate=16),
                                   7).cells == 4
  j3 = plan.make_tuned_pass_plan(common.stencil('jacobi3d', iterate=32), 1)
  assert (j3.rows, j3.cy, j3.pack, j3.skew) == (8, 4, 1, 0)
  j3 = plan.make_tuned_pass_plan(common.stencil('jacobi3d', iterate=32), 2)
  assert (j3.rows, j3.cy, j3.valid) == (16, 4, (120, 12))
  d3 = plan.make_tuned_pass_plan(common.stencil('denoise3d'), 1)
  assert (d3.rows, d3.cy) == (24, 1)  # 11 nodes: patches would spill
  # explicit options win
  j3 = plan.make_tuned_pass_plan(common.stencil('jacobi3d', iterate=32), 2,
                                 {'rows': 32, 'cy': 2, 'pack': True})
  assert (j3.rows, j3.cy, j3.pack) == (32, 2, 2)
  text = emit.emit_program(common.stencil('jacobi3d', iterate=32), 1)
  assert 'kCy = 4' in text and 'kWarps = 2' in text and 'kUnroll = 3' in text


def test_patch_rows_export_only_their_edges():
  p = plan.make_pass_plan(common.stencil('jacobi3d', iterate=32), time_block=2,
                          rows=16, cy=4, pack=False)
  t1, mid, out = p.nodes
  assert mid.smem_depth == 2 and mid.smem_reach == 1
  # with patches the dimension-1 neighbours inside the patch come from the
  # register window, which therefore covers their (skewed) distance too
  assert mid.ring == 3 and t1.ring == 3
  with pytest.raises(util.SemanticError):
    plan.make_pass_plan(common.stencil('jacobi3d'), rows=10, cy=4)


def test_3d_tiles_fit_shared_memory_and_the_thread_limit():
  """Found by tests/test_random_programs.py seed 13 on B200: two half inputs
  and two exported stages with a dimension-1 halo of 5 rows asked for a 30-row
  tile = 276 KB of shared memory.  The tuned plan lowers the tile until it
  fits; explicit oversize requests fail at compile time, not at launch."""
  from tests import random_programs
  text, _, _ = random_programs.program(13)
  st = sodac.compile_source(text)
  tuned = plan.make_tuned_pass_plan(st, 1)
  geometry = plan.smem_geometry_3d(tuned)
  assert geometry['bytes'] <= plan.SMEM_LIMIT_BYTES
  assert tuned.rows // tuned.cy * 32 <= plan.MAX_CTA_THREADS
  assert tuned.rows > tuned.halo_lo[1] + tuned.halo_hi[1]
  with pytest.raises(util.SemanticError, match='shared memory'):
    emit.emit_program(st, options={'rows': 32})
  with pytest.raises(util.SemanticError, match='threads per CTA'):
    emit.emit_program(common.stencil('jacobi3d'), options={'rows': 40, 'cy': 1})
  # the programs of tests/src keep their measured tiles
  jacobi = plan.make_tuned_pass_plan(common.stencil('jacobi3d', iterate=4), 2)
  assert (jacobi.rows, jacobi.cy) == (16, 4)
  denoise = plan.make_tuned_pass_plan(common.stencil('denoise3d'), 1)
  assert (denoise.rows, denoise.cy) == (24, 1)
