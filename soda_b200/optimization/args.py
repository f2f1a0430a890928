"""Command-line flags of the SODA optimisation passes.

Same flags and the same ``get_kwargs`` contract as the reference
(reference: src/soda/optimization/args.py:5-31): the returned dict becomes
``Stencil(optimizations=...)``.
"""
import argparse
from typing import Dict

CR_CHOICES = ('yes', 'no', 'greedy', 'optimal', 'glore', 'built-in',
              'built-in:greedy', 'built-in:optimal')


def add_arguments(parser) -> None:
  parser.add_argument('--inline',
                      type=str,
                      metavar='(yes|no)',
                      dest='inline',
                      nargs='?',
                      const='yes',
                      default='no',
                      help='inline locals that are referenced exactly once')
  parser.add_argument('--computation-reuse',
                      type=str,
                      metavar='(%s)' % '|'.join(CR_CHOICES),
                      dest='computation_reuse',
                      nargs='?',
                      const='yes',
                      default='no',
                      help='enable computation reuse or not')


def get_kwargs(args: argparse.Namespace) -> Dict[str, str]:
  optimizations = {}
  if args.computation_reuse != 'no':
    optimizations['computation-reuse'] = args.computation_reuse
  if args.inline != 'no':
    optimizations['inline'] = args.inline
  return optimizations
