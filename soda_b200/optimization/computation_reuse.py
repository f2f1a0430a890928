"""Computation reuse (soda-cr).  Placeholder pass-through until the scheduler
lands; see SURVEY.md section 8(a) row a7."""


def computation_reuse(stencil):
  return stencil
