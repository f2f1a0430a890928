"""Command-line flags of the SODA optimisation passes.

One table row per pass.  The flag spellings, their accepted values and the
``get_kwargs`` contract are the reference's (reference:
src/soda/optimization/args.py:5-31): a flag given without a value means
``yes``, an absent flag means ``no``, and every pass that is not ``no`` lands in
the dict that becomes ``Stencil(optimizations=...)`` under the flag's own
spelling (without the dashes in front).
"""
import argparse
from typing import Dict, NamedTuple, Tuple


class _Pass(NamedTuple):
  key: str  # key in Stencil.optimizations, also the flag without '--'
  values: Tuple[str, ...]
  text: str

  @property
  def attribute(self) -> str:
    return self.key.replace('-', '_')


PASSES = (
    _Pass('computation-reuse',
          ('yes', 'no', 'greedy', 'optimal', 'glore', 'built-in',
           'built-in:greedy', 'built-in:optimal'),
          'enable computation reuse or not'),
    _Pass('inline', ('yes', 'no'),
          'inline locals that are referenced exactly once'),
)


def add_arguments(parser) -> None:
  for item in PASSES:
    parser.add_argument('--' + item.key, dest=item.attribute, type=str,
                        nargs='?', const='yes', default='no',
                        metavar='(%s)' % '|'.join(item.values), help=item.text)


def get_kwargs(args: argparse.Namespace) -> Dict[str, str]:
  chosen = ((item.key, getattr(args, item.attribute)) for item in PASSES)
  return {key: value for key, value in chosen if value != 'no'}
