"""The planner's throughput model (soda_b200/codegen/cuda/model.py) against
the sweep measured on a B200 (profiles/r02_time_block_sweep.jsonl, written by
tools/tb_sweep.py): for every swept program the time block the model picks
must be within a few per cent of the measured best, and its absolute
prediction within a factor the model can be trusted for when it reports
ceilings (bench.py ``other_configs``)."""
import collections
import json
import os

import pytest

from soda_b200.codegen.cuda import model
from tests import common

SWEEP = os.path.join(common.ROOT, 'profiles', 'r02_time_block_sweep.jsonl')


def load_sweep():
  table = collections.defaultdict(dict)
  with open(SWEEP) as fp:
    for line in fp:
      row = json.loads(line)
      if 'gcells' not in row or row['options']:
        continue  # only the planner's default shape of each time block
      table[(row['program'], tuple(row['extent']))][row['tb']] = row['gcells']
  return table


@pytest.mark.parametrize('key', sorted(load_sweep()))
def test_model_picks_a_time_block_near_the_measured_best(key):
  name, extent = key
  measured = load_sweep()[key]
  st = common.stencil(name, iterate=4 * max(measured))
  choice = model.choose_time_block(st, None, list(extent),
                                   limit=max(measured))
  assert choice in measured
  best = max(measured.values())
  # the planner prefers a smaller block within 12 % (modelled) of the best
  assert measured[choice] >= (1.0 - model.NEAR_TIE - 0.04) * best, \
      (choice, measured)


@pytest.mark.parametrize('key', sorted(load_sweep()))
def test_model_predictions_are_in_range(key):
  name, extent = key
  for tb, gcells in load_sweep()[key].items():
    st = common.stencil(name, iterate=4 * tb)
    est = model.estimate_pass(st, tb, None, list(extent))
    assert est is not None
    assert 0.6 < est['gcells'] / gcells < 1.5, (tb, est['gcells'], gcells)


def test_register_budget_rules_out_what_cannot_be_built():
  st = common.stencil('jacobi2d', iterate=64)
  wide = model.estimate_pass(st, 8)
  assert wide['cells'] == 8 and wide['window_registers'] == 200
  assert model.estimate_pass(st, 9)['cells'] == 4  # 224 window registers
  assert model.estimate_pass(st, 12, {'cells': 8}) is None
