"""Python host for a compiled SODA program: a thin ctypes layer over the C ABI
declared in include/soda_cuda.h.

This is the host the north star asks for ("a Python host calls CUDA through a
thin ctypes C-ABI layer").  It mirrors what the reference's generated C++ host
does around the kernel (reference: src/soda/codegen/frt/host.py:62-431): take
the caller's arrays, run all ``iterate`` iterations, write only the valid
interior of every output.  There is no CPU path: loading fails loudly when the
library is missing, and every call reports the library's own error string.
"""
import ctypes
import os
from typing import Dict, Optional, Sequence

import numpy as np

MAX_DIM = 3
MAX_TENSORS = 8

DTYPES = ('uint8', 'int8', 'uint16', 'int16', 'uint32', 'int32', 'uint64',
          'int64', 'float32', 'float64', 'float16')

EXPORTED_SYMBOLS = (
    'soda_cuda_info',
    'soda_cuda_get_pass_info',
    'soda_cuda_last_error',
    'soda_cuda_run_host',
    'soda_cuda_plan_create',
    'soda_cuda_plan_destroy',
    'soda_cuda_plan_run_host',
    'soda_cuda_plan_run_device',
    'soda_cuda_run_pass',
    'soda_cuda_launch_count',
    'soda_cuda_set_param',
    # multi-GPU slabs
    'soda_cuda_nccl_unique_id',
    'soda_cuda_slab_create',
    'soda_cuda_slab_destroy',
    'soda_cuda_slab_get_info',
    'soda_cuda_slab_buffers',
    'soda_cuda_slab_run',
    'soda_cuda_slab_exchange_inputs',
    'soda_cuda_slab_run_host',
    'soda_cuda_multi_run_host',
    'soda_cuda_host_alloc',
    'soda_cuda_host_free',
)


class SodaCudaError(RuntimeError):

  def __init__(self, status: int, message: str):
    super().__init__('soda_cuda status %d: %s' % (status, message))
    self.status = status


class Opts(ctypes.Structure):
  _fields_ = [
      ('struct_size', ctypes.c_int32),
      ('device', ctypes.c_int32),
      ('stream', ctypes.c_void_p),
      ('segment', ctypes.c_int32),
      ('reserved', ctypes.c_int32 * 5),
  ]


class PassInfo(ctypes.Structure):
  _fields_ = [
      ('time_block', ctypes.c_int32),
      ('reach_lo', ctypes.c_int32 * MAX_DIM),
      ('reach_hi', ctypes.c_int32 * MAX_DIM),
      ('threads_per_cta', ctypes.c_int32),
      ('smem_bytes', ctypes.c_int32),
      ('cells_per_lane', ctypes.c_int32),
      ('strip_cells', ctypes.c_int32),
      ('valid_cells', ctypes.c_int32 * 2),
  ]


class ProgramInfo(ctypes.Structure):
  _fields_ = [
      ('app_name', ctypes.c_char_p),
      ('soda_source', ctypes.c_char_p),
      ('dim', ctypes.c_int32),
      ('iterate', ctypes.c_int32),
      ('num_inputs', ctypes.c_int32),
      ('num_outputs', ctypes.c_int32),
      ('input_names', ctypes.c_char_p * MAX_TENSORS),
      ('output_names', ctypes.c_char_p * MAX_TENSORS),
      ('input_dtypes', ctypes.c_int32 * MAX_TENSORS),
      ('output_dtypes', ctypes.c_int32 * MAX_TENSORS),
      ('final_lo', (ctypes.c_int32 * MAX_DIM) * MAX_TENSORS),
      ('final_hi', (ctypes.c_int32 * MAX_DIM) * MAX_TENSORS),
      ('num_passes', ctypes.c_int32),
      ('strict_fp', ctypes.c_int32),
      ('algorithmic_bytes_per_cell_per_pass', ctypes.c_int32),
      ('num_params', ctypes.c_int32),
      ('param_names', ctypes.c_char_p * MAX_TENSORS),
      ('param_dtypes', ctypes.c_int32 * MAX_TENSORS),
      ('param_elems', ctypes.c_int32 * MAX_TENSORS),
      ('source_dim', ctypes.c_int32),
  ]


def make_opts(device: int = -1, stream: int = 0, segment: int = 0,
              host_chunks: int = 0, gpus: int = 0) -> Opts:
  """``host_chunks``: chunk count of the copy/compute pipeline used for host
  arrays (0 = automatic, 1 = no pipelining).  ``gpus`` > 1: host-array calls
  split the grid over that many devices of this process."""
  opts = Opts()
  opts.struct_size = ctypes.sizeof(Opts)
  opts.device = device
  opts.stream = stream or None
  opts.segment = segment
  opts.reserved[0] = host_chunks
  opts.reserved[1] = gpus
  return opts


Pitch2 = ctypes.c_int64 * 2
Box = ctypes.c_int32 * MAX_DIM


class CudaProgram:
  """A loaded ``libsoda_<app>.so``."""

  def __init__(self, lib_path: str):
    if not os.path.exists(lib_path):
      raise FileNotFoundError(
          'compiled SODA program %s not found; build it with sodac '
          '--cuda-lib (there is no CPU fallback)' % lib_path)
    self.lib_path = lib_path
    self.lib = ctypes.CDLL(lib_path)
    for name in EXPORTED_SYMBOLS:
      if not hasattr(self.lib, name):
        raise SodaCudaError(-1, '%s does not export %s' % (lib_path, name))
    self.lib.soda_cuda_last_error.restype = ctypes.c_char_p
    self.lib.soda_cuda_launch_count.restype = ctypes.c_int64
    info = ProgramInfo()
    self._check(self.lib.soda_cuda_info(ctypes.byref(info)))
    self.info = info
    self.app_name = info.app_name.decode()
    self.soda_source = info.soda_source.decode()
    self.dim = info.dim
    # a 1-D program runs lifted to an N x 1 grid (include/soda_cuda.h)
    self.source_dim = info.source_dim or info.dim
    self.iterate = info.iterate
    self.input_names = [info.input_names[i].decode()
                        for i in range(info.num_inputs)]
    self.output_names = [info.output_names[i].decode()
                         for i in range(info.num_outputs)]
    self.input_dtypes = [np.dtype(DTYPES[info.input_dtypes[i]])
                         for i in range(info.num_inputs)]
    self.output_dtypes = [np.dtype(DTYPES[info.output_dtypes[i]])
                          for i in range(info.num_outputs)]
    self.param_names = [info.param_names[i].decode()
                        for i in range(info.num_params)]
    self.param_dtypes = [np.dtype(DTYPES[info.param_dtypes[i]])
                         for i in range(info.num_params)]
    self.param_elems = [int(info.param_elems[i])
                        for i in range(info.num_params)]
    self.num_passes = info.num_passes
    self.bytes_per_cell_per_pass = info.algorithmic_bytes_per_cell_per_pass
    self.app_entry = getattr(self.lib, 'soda_cuda_' + self.app_name)

  # -- helpers ---------------------------------------------------------------
  def _check(self, status: int) -> None:
    if status != 0:
      raise SodaCudaError(status, self.lib.soda_cuda_last_error().decode())

  def pass_info(self, index: int) -> PassInfo:
    info = PassInfo()
    self._check(self.lib.soda_cuda_get_pass_info(index, ctypes.byref(info)))
    return info

  def launch_count(self) -> int:
    return int(self.lib.soda_cuda_launch_count())

  def valid_box(self, output: int, extent: Sequence[int]):
    """[(lo, hi)] per dimension of the cells of ``output`` that get written."""
    return [(self.info.final_lo[output][d],
             extent[d] - self.info.final_hi[output][d])
            for d in range(self.dim)]

  def _lifted(self, array: np.ndarray) -> np.ndarray:
    """1-D arrays of a 1-D program as the N x 1 grid the library runs."""
    if self.source_dim == 1 and array.ndim == 1:
      return array.reshape(1, -1)
    return array

  def _extent_of(self, array: np.ndarray):
    if array.ndim != self.dim:
      raise ValueError('%d-D array for a %d-D program' % (array.ndim, self.dim))
    return tuple(array.shape[::-1])

  @staticmethod
  def _strides_of(array: np.ndarray):
    if any(s % array.itemsize for s in array.strides):
      raise ValueError('array strides must be multiples of the element size')
    return tuple(s // array.itemsize for s in array.strides[::-1])

  # -- params ------------------------------------------------------------------
  def _param_array(self, index: int, values) -> np.ndarray:
    array = np.ascontiguousarray(values, dtype=self.param_dtypes[index])
    if array.size != self.param_elems[index]:
      raise ValueError('param %s has %d elements, got %d' %
                       (self.param_names[index], self.param_elems[index],
                        array.size))
    return array

  def set_params(self, params: Dict[str, np.ndarray],
                 opts: Optional[Opts] = None) -> None:
    """Values of the program's ``param`` arrays (used by every later launch
    on the device of ``opts``)."""
    for index, name in enumerate(self.param_names):
      array = self._param_array(index, params[name])
      self._check(self.lib.soda_cuda_set_param(
          index, ctypes.c_void_p(array.ctypes.data),
          ctypes.byref(opts) if opts is not None else None))

  # -- host arrays -------------------------------------------------------------
  def run_host(self,
               inputs: Dict[str, np.ndarray],
               outputs: Optional[Dict[str, np.ndarray]] = None,
               opts: Optional[Opts] = None,
               use_app_entry: bool = True,
               params: Optional[Dict[str, np.ndarray]] = None
               ) -> Dict[str, np.ndarray]:
    """Runs the program on NumPy arrays (shape ``extent[::-1]``, dimension 0
    contiguous).  Only the valid interior of each output is written; pass
    ``outputs`` to see that the rest is left untouched.  ``params``: values of
    the program's ``param`` arrays (required when it declares any and they were
    not set with ``set_params``)."""
    if self.param_names and params is None and use_app_entry:
      raise ValueError('program %s needs params %s' %
                       (self.app_name, self.param_names))
    flat = self.source_dim == 1 and inputs[self.input_names[0]].ndim == 1
    ins = []
    for name, dtype in zip(self.input_names, self.input_dtypes):
      array = self._lifted(inputs[name])
      if array.dtype != dtype:
        raise TypeError('input %s must be %s, got %s' %
                        (name, dtype, array.dtype))
      ins.append(array)
    extent = self._extent_of(ins[0])
    if outputs is None:
      outputs = {
          name: np.zeros(extent[::-1][1:] if flat else extent[::-1],
                         dtype=dtype)
          for name, dtype in zip(self.output_names, self.output_dtypes)
      }
    outs = []
    for name, dtype in zip(self.output_names, self.output_dtypes):
      array = self._lifted(outputs[name])
      if array.dtype != dtype or self._extent_of(array) != extent:
        raise TypeError('output %s must be %s of extent %s' %
                        (name, dtype, extent))
      if not array.flags.writeable:
        raise ValueError('output %s is read-only' % name)
      outs.append(array)
    for array in ins:
      if self._extent_of(array) != extent:
        raise ValueError('all tensors must share one extent')
    c_extent = (ctypes.c_int32 * self.dim)(*extent)
    zeros = (ctypes.c_int32 * self.dim)(*([0] * self.dim))
    strides = [(ctypes.c_int32 * self.dim)(*self._strides_of(a))
               for a in ins + outs]
    opts_ref = ctypes.byref(opts) if opts is not None else None
    if use_app_entry:
      # the program-named entry point: (ptr, extent, stride, min) per tensor,
      # in the dimensions of the program as written
      args = []
      n = self.source_dim
      for array, stride in zip(ins + outs, strides):
        args += [ctypes.c_void_p(array.ctypes.data),
                 (ctypes.c_int32 * n)(*extent[:n]),
                 (ctypes.c_int32 * n)(*list(stride)[:n]),
                 (ctypes.c_int32 * n)(*([0] * n))]
      keep = []
      for index, name in enumerate(self.param_names):
        array = self._param_array(index, params[name])
        keep.append(array)
        shape = list(np.shape(params[name]))[::-1] or [array.size]
        p_extent = (ctypes.c_int32 * len(shape))(*shape)
        p_zero = (ctypes.c_int32 * len(shape))(*([0] * len(shape)))
        p_stride = (ctypes.c_int32 * len(shape))(
            *[int(np.prod(shape[:d])) for d in range(len(shape))])
        args += [ctypes.c_void_p(array.ctypes.data), p_extent, p_stride, p_zero]
      args.append(opts_ref)
      self._check(self.app_entry(*args))
    else:
      if params is not None:
        self.set_params(params, opts)
      n_in, n_out = len(ins), len(outs)
      in_ptrs = (ctypes.c_void_p * n_in)(*[a.ctypes.data for a in ins])
      out_ptrs = (ctypes.c_void_p * n_out)(*[a.ctypes.data for a in outs])
      stride_ptr = ctypes.POINTER(ctypes.c_int32)
      in_strides = (stride_ptr * n_in)(
          *[ctypes.cast(s, stride_ptr) for s in strides[:n_in]])
      out_strides = (stride_ptr * n_out)(
          *[ctypes.cast(s, stride_ptr) for s in strides[n_in:]])
      self._check(
          self.lib.soda_cuda_run_host(in_ptrs, in_strides, out_ptrs,
                                      out_strides, c_extent, opts_ref))
    return outputs

  # -- plans -------------------------------------------------------------------
  def create_plan(self, extent: Sequence[int],
                  opts: Optional[Opts] = None) -> 'Plan':
    return Plan(self, extent, opts)

  def run_pass(self, pass_index: int, extent: Sequence[int],
               d_in: Sequence[int], in_pitches, d_out: Sequence[int],
               out_pitches, box_lo=None, box_hi=None,
               opts: Optional[Opts] = None) -> None:
    """One pass on raw device pointers (ints); pitches are (row, plane) element
    pitches per tensor; boxes are per-output lists of per-dimension bounds."""
    n_in, n_out = len(d_in), len(d_out)
    c_extent = (ctypes.c_int32 * self.dim)(*extent)
    in_ptrs = (ctypes.c_void_p * n_in)(*d_in)
    out_ptrs = (ctypes.c_void_p * n_out)(*d_out)
    in_p = (Pitch2 * n_in)(*[Pitch2(*p) for p in in_pitches])
    out_p = (Pitch2 * n_out)(*[Pitch2(*p) for p in out_pitches])
    lo = hi = None
    if box_lo is not None:
      pad = lambda b: list(b) + [0] * (MAX_DIM - len(b))
      lo = (Box * n_out)(*[Box(*pad(b)) for b in box_lo])
      hi = (Box * n_out)(*[Box(*pad(b)) for b in box_hi])
    self._check(
        self.lib.soda_cuda_run_pass(pass_index, c_extent, in_ptrs, in_p,
                                    out_ptrs, out_p, lo, hi,
                                    ctypes.byref(opts) if opts else None))


class HostBuffer:
  """Page-locked host array on the NUMA node of a device
  (``soda_cuda_host_alloc``): the staging memory of the host entry points."""

  def __init__(self, program: 'CudaProgram', shape: Sequence[int], dtype,
               device: int = -1):
    self.program = program
    self.dtype = np.dtype(dtype)
    self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
    self.ptr = ctypes.c_void_p()
    program._check(program.lib.soda_cuda_host_alloc(
        ctypes.byref(self.ptr), ctypes.c_int64(self.nbytes), device))
    raw = (ctypes.c_char * self.nbytes).from_address(self.ptr.value)
    self.array = np.frombuffer(raw, dtype=self.dtype).reshape(tuple(shape))

  def close(self) -> None:
    if self.ptr:
      self.array = None
      self.program.lib.soda_cuda_host_free(self.ptr,
                                           ctypes.c_int64(self.nbytes))
      self.ptr = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


class Plan:
  """Device-side scratch for one extent; reusable across calls."""

  def __init__(self, program: CudaProgram, extent: Sequence[int],
               opts: Optional[Opts] = None):
    self.program = program
    self.extent = tuple(extent)
    self.handle = ctypes.c_void_p()
    c_extent = (ctypes.c_int32 * program.dim)(*extent)
    program._check(
        program.lib.soda_cuda_plan_create(
            c_extent, ctypes.byref(opts) if opts is not None else None,
            ctypes.byref(self.handle)))

  def run_host(self, inputs: Dict[str, np.ndarray],
               outputs: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    prog = self.program
    ins = [inputs[name] for name in prog.input_names]
    outs = [outputs[name] for name in prog.output_names]
    for array, dtype in zip(ins + outs, prog.input_dtypes + prog.output_dtypes):
      if array.dtype != dtype or tuple(array.shape[::-1]) != self.extent:
        raise TypeError('tensor must be %s of extent %s' % (dtype, self.extent))
    stride_ptr = ctypes.POINTER(ctypes.c_int32)
    strides = [(ctypes.c_int32 * prog.dim)(*prog._strides_of(a))
               for a in ins + outs]
    n_in, n_out = len(ins), len(outs)
    in_ptrs = (ctypes.c_void_p * n_in)(*[a.ctypes.data for a in ins])
    out_ptrs = (ctypes.c_void_p * n_out)(*[a.ctypes.data for a in outs])
    in_strides = (stride_ptr * n_in)(
        *[ctypes.cast(s, stride_ptr) for s in strides[:n_in]])
    out_strides = (stride_ptr * n_out)(
        *[ctypes.cast(s, stride_ptr) for s in strides[n_in:]])
    prog._check(
        prog.lib.soda_cuda_plan_run_host(self.handle, in_ptrs, in_strides,
                                         out_ptrs, out_strides))
    return outputs

  def run_device(self, d_in: Sequence[int], in_pitches, d_out: Sequence[int],
                 out_pitches) -> None:
    """All iterations on raw device pointers, asynchronous on the plan's
    stream."""
    prog = self.program
    n_in, n_out = len(d_in), len(d_out)
    in_ptrs = (ctypes.c_void_p * n_in)(*d_in)
    out_ptrs = (ctypes.c_void_p * n_out)(*d_out)
    in_p = (Pitch2 * n_in)(*[Pitch2(*p) for p in in_pitches])
    out_p = (Pitch2 * n_out)(*[Pitch2(*p) for p in out_pitches])
    prog._check(
        prog.lib.soda_cuda_plan_run_device(self.handle, in_ptrs, in_p, out_ptrs,
                                           out_p))

  def close(self) -> None:
    if self.handle:
      self.program.lib.soda_cuda_plan_destroy(self.handle)
      self.handle = ctypes.c_void_p()

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass
