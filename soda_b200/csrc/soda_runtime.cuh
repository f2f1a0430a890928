// Host side of a compiled SODA program: implements the C ABI declared in
// include/soda_cuda.h on top of the kernel templates in soda_stream.cuh.
//
// Included at the end of every generated program file, after the generated
// code has defined
//     static const soda::rt::ProgramDesc& soda_program();
// Nothing here depends on the program; everything program-specific arrives
// through ProgramDesc (metadata) and PassImpl::launch (instantiated template).
#pragma once

#ifndef SODA_EMU
#include <cuda.h>
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "soda_cuda.h"
#include "soda_stream.cuh"

namespace soda {
namespace rt {

constexpr int kMaxT = SODA_CUDA_MAX_TENSORS;
constexpr int kMaxD = SODA_CUDA_MAX_DIM;
constexpr int kNumSms = 148;  // B200

// ---- errors -----------------------------------------------------------------
inline std::string& last_error() {
  static thread_local std::string message;
  return message;
}

inline int fail(int status, const std::string& message) {
  last_error() = message;
  return status;
}

#define SODA_CUDA_CHECK(expr)                                              \
  do {                                                                     \
    cudaError_t soda_err_ = (expr);                                        \
    if (soda_err_ != cudaSuccess) {                                        \
      return ::soda::rt::fail(                                             \
          soda_err_ == cudaErrorMemoryAllocation ? SODA_CUDA_OUT_OF_MEMORY \
                                                 : SODA_CUDA_CUDA_ERROR,   \
          std::string(#expr) + ": " + cudaGetErrorString(soda_err_));      \
    }                                                                      \
  } while (0)

inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> counter{0};
  return counter;
}

// ---- one pass ---------------------------------------------------------------
struct PassArgs {
  int extent[kMaxD];
  const void* in[kMaxT];
  long long in_pitch[kMaxT][2];  // elements: between rows, between planes
  void* out[kMaxT];
  long long out_pitch[kMaxT][2];
  int box_lo[kMaxT][kMaxD];  // cells each output may be written in
  int box_hi[kMaxT][kMaxD];
  int segment;  // output slices per CTA along the streamed dim, 0 = auto
  cudaStream_t stream;
};

struct PassImpl {
  soda_cuda_pass_info info;
  int (*launch)(const PassArgs&);
  int warmup;  // slices a segment computes before its first stored slice
};

struct ProgramDesc {
  soda_cuda_program_info info;
  int in_elem_bytes[kMaxT];
  int out_elem_bytes[kMaxT];
  const PassImpl* impls;
  int num_impls;
  const int* schedule;  // impl index of every pass
  // `param` arrays: constant-memory symbols of the generated file
  const void* param_symbol[kMaxT];
  int param_bytes[kMaxT];
};

// ---- TMA descriptors ----------------------------------------------------------
template <typename T>
constexpr CUtensorMapDataType tma_dtype() {
  if (std::is_same<T, float>::value) return CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (std::is_same<T, double>::value) return CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  if (sizeof(T) == 1) return CU_TENSOR_MAP_DATA_TYPE_UINT8;
  if (sizeof(T) == 2) return CU_TENSOR_MAP_DATA_TYPE_UINT16;
  if (sizeof(T) == 4)
    return std::is_signed<T>::value ? CU_TENSOR_MAP_DATA_TYPE_INT32
                                    : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  return std::is_signed<T>::value ? CU_TENSOR_MAP_DATA_TYPE_INT64
                                  : CU_TENSOR_MAP_DATA_TYPE_UINT64;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t,
                                   void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult query;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault,
                                &query) != cudaSuccess ||
        query != cudaDriverEntryPointSuccess) {
      ptr = nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

inline CUtensorMapL2promotion l2_promotion() {
#ifndef SODA_EMU
  const char* env = getenv("SODA_CUDA_L2_PROMOTION");
  if (env != nullptr) {
    switch (atoi(env)) {
      case 0: return CU_TENSOR_MAP_L2_PROMOTION_NONE;
      case 64: return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
      case 128: return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
      default: break;
    }
  }
#endif
  return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// Tiled map over a dense-in-dim-0 tensor: extent[d] cells, pitch[d-1] elements
// between indices of dimension d; box[d] cells per load.  Out-of-range cells
// read as zero.
template <typename T>
int make_tensor_map(CUtensorMap* map, const void* base, int rank,
                    const int* extent, const long long* pitch,
                    const int* box) {
  EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr)
    return fail(SODA_CUDA_CUDA_ERROR,
                "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[kMaxD];
  cuuint64_t strides[kMaxD];
  cuuint32_t box_dims[kMaxD];
  cuuint32_t elem_strides[kMaxD];
  for (int d = 0; d < rank; ++d) {
    dims[d] = static_cast<cuuint64_t>(extent[d]);
    box_dims[d] = static_cast<cuuint32_t>(box[d]);
    elem_strides[d] = 1;
    if (d > 0) strides[d - 1] = static_cast<cuuint64_t>(pitch[d - 1]) * sizeof(T);
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0)
    return fail(SODA_CUDA_BAD_ARGUMENT,
                "device arrays must be 16-byte aligned for TMA");
  for (int d = 0; d + 1 < rank; ++d) {
    if (strides[d] % 16 != 0)
      return fail(SODA_CUDA_BAD_ARGUMENT,
                  "device array pitches must be multiples of 16 bytes for TMA");
  }
  CUresult res = encode(map, tma_dtype<T>(), static_cast<cuuint32_t>(rank),
                        const_cast<void*>(base), dims, strides, box_dims,
                        elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(),
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS)
    return fail(SODA_CUDA_CUDA_ERROR, "cuTensorMapEncodeTiled failed with code " +
                                          std::to_string(static_cast<int>(res)));
  return SODA_CUDA_OK;
}

template <class Prog, int M = 0>
int make_input_maps(CUtensorMap* maps, const PassArgs& a, const int* box) {
  if constexpr (M < Prog::kNumInputs) {
    using T = typename Prog::template T<M>;
    int status = make_tensor_map<T>(&maps[M], a.in[M], Prog::kDim, a.extent,
                                    a.in_pitch[M], box);
    if (status != SODA_CUDA_OK) return status;
    return make_input_maps<Prog, M + 1>(maps, a, box);
  } else {
    return SODA_CUDA_OK;
  }
}

template <class Prog, int O = 0>
bool outputs_vector_aligned(const PassArgs& a) {
  if constexpr (O < Prog::kNumOutputs) {
    constexpr int kNode = Prog::kOutputNode[O];
    using T = typename Prog::template T<kNode>;
    constexpr size_t kVec = vec_piece_bytes(sizeof(T) * Prog::kCells);
    bool ok = reinterpret_cast<uintptr_t>(a.out[O]) % kVec == 0 &&
              (a.out_pitch[O][0] * sizeof(T)) % kVec == 0 &&
              (Prog::kDim < 3 || (a.out_pitch[O][1] * sizeof(T)) % kVec == 0);
    return ok && outputs_vector_aligned<Prog, O + 1>(a);
  } else {
    return true;
  }
}

inline int floor_to(int value, int multiple) {
  int q = value / multiple;
  if (value % multiple != 0 && value < 0) --q;
  return q * multiple;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Shape of the most recent launch on this thread: CTAs per set of segments
// (tiles across the non-streamed dimensions) and CTAs the GPU holds at once.
struct LaunchShape {
  int tiles;
  int slots;
};
inline LaunchShape& last_launch_shape() {
  static thread_local LaunchShape shape = {0, 0};
  return shape;
}

// SODA_CUDA_WAVE_RULE=0 (experiments): choose_segment() without its
// fill-one-wave-first clause.
inline bool wave_rule_enabled() {
  static const bool enabled = [] {
    const char* env = getenv("SODA_CUDA_WAVE_RULE");
    return env == nullptr || env[0] != '0';
  }();
  return enabled;
}

// Segments along the streamed dimension, by rule: enough CTAs for about four
// waves, but long enough that the warm-up slices stay a small fraction.  Large
// grids replace the rule by a measurement (launch_tuned below).
inline int choose_segment(int slices, int ctas_per_slice_set, int warmup,
                          int ctas_per_sm, int requested) {
  if (requested > 0) return requested < slices ? requested : slices;
  const int min_segment = warmup * 16 > 64 ? warmup * 16 : 64;
  const long long target_ctas = 4LL * kNumSms * ctas_per_sm;
  long long segments = target_ctas / (ctas_per_slice_set > 0 ? ctas_per_slice_set : 1);
  if (segments < 1) segments = 1;
  long long max_segments = slices / min_segment;
  if (max_segments < 1) max_segments = 1;
  if (segments > max_segments) segments = max_segments;
  // A window too short for that many long segments would leave SMs without a
  // CTA (a 16384 x 640 window of the bench kernel: 4 segments x 34 CTAs on 148
  // SMs, 0.94 ms for 11 passes against 0.42 ms with 17 segments): fill one
  // wave first, with segments down to 4x the warm-up.
  const long long one_wave = static_cast<long long>(kNumSms) * ctas_per_sm;
  const long long per_set = ctas_per_slice_set > 0 ? ctas_per_slice_set : 1;
  if (segments * per_set < one_wave && wave_rule_enabled()) {
    const int short_segment = warmup * 4 > 16 ? warmup * 4 : 16;
    long long relaxed = slices / short_segment;
    const long long wanted = (one_wave + per_set - 1) / per_set;
    if (relaxed > wanted) relaxed = wanted;
    if (relaxed > segments) segments = relaxed;
  }
  return ceil_div(slices, static_cast<int>(segments));
}

// The dynamic shared-memory attribute and the occupancy of a kernel are per
// device: one process may drive several (soda_cuda_opts.device, the slab and
// multi-device entry points).  Set / queried once per (kernel, device).
constexpr int kMaxDevices = 64;
template <class Kernel>
int kernel_setup(Kernel kernel, int threads, int smem_bytes, int* ctas_per_sm) {
  static std::mutex mutex;
  static bool done[kMaxDevices] = {};
  static int occupancy[kMaxDevices] = {};
  int device = 0;
  SODA_CUDA_CHECK(cudaGetDevice(&device));
  if (device < 0 || device >= kMaxDevices)
    return fail(SODA_CUDA_UNSUPPORTED, "device ordinal out of range");
  std::lock_guard<std::mutex> lock(mutex);
  if (!done[device]) {
    SODA_CUDA_CHECK(cudaFuncSetAttribute(
        kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int blocks = 1;
    SODA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &blocks, kernel, threads, smem_bytes));
    occupancy[device] = blocks < 1 ? 1 : blocks;
    done[device] = true;
  }
  *ctas_per_sm = occupancy[device];
  return SODA_CUDA_OK;
}

// Grids of at most this many strips run as one-warp CTAs (Smem2D);
// SODA_CUDA_NARROW_STRIPS overrides the threshold (0: never; experiments and
// the tests that want the several-warp CTAs on small grids).
inline int narrow_grid_strips() {
  static const int value = [] {
    const char* env = getenv("SODA_CUDA_NARROW_STRIPS");
    return env != nullptr ? atoi(env) : 32;
  }();
  return value;
}

template <class Prog, int kWarps>
int launch_pass_2d_as(const PassArgs& a) {
  using S = Smem2D<Prog, kWarps>;
  auto kernel = soda_stream2d_kernel<Prog, kWarps>;
  int ctas_per_sm = 1;
  {
    int status = kernel_setup(kernel, kWarps * 32, S::kBytes, &ctas_per_sm);
    if (status != SODA_CUDA_OK) return status;
  }

  Params2D<Prog> p;
  memset(&p, 0, sizeof(p));
  const int box[2] = {S::kBox0, Prog::kChunk};  // S::kBoxes of them per strip
  int status = make_input_maps<Prog>(p.in_map, a, box);
  if (status != SODA_CUDA_OK) return status;

  int x_lo = a.extent[0], x_hi = 0, row_lo = a.extent[1], row_hi = 0;
  for (int o = 0; o < Prog::kNumOutputs; ++o) {
    p.out[o] = a.out[o];
    p.out_pitch[o] = a.out_pitch[o][0];
    for (int d = 0; d < 2; ++d) {
      p.box_lo[o][d] = a.box_lo[o][d];
      p.box_hi[o][d] = a.box_hi[o][d];
    }
    if (a.box_lo[o][0] < x_lo) x_lo = a.box_lo[o][0];
    if (a.box_hi[o][0] > x_hi) x_hi = a.box_hi[o][0];
    if (a.box_lo[o][1] < row_lo) row_lo = a.box_lo[o][1];
    if (a.box_hi[o][1] > row_hi) row_hi = a.box_hi[o][1];
  }
  if (x_hi <= x_lo || row_hi <= row_lo) return SODA_CUDA_OK;  // nothing valid

  p.x_origin = floor_to(x_lo, Prog::kAlign0);
  p.num_strips = ceil_div(x_hi - p.x_origin, Prog::kValid0);
  p.row_lo = row_lo;
  p.row_hi = row_hi;
  const int ctas_x = ceil_div(p.num_strips, kWarps);
  p.seg_rows = choose_segment(row_hi - row_lo, ctas_x,
                              Prog::kMaxLag - Prog::kLoS, ctas_per_sm, a.segment);
  p.vec_ok = outputs_vector_aligned<Prog>(a) ? 1 : 0;
  dim3 grid(ctas_x, ceil_div(row_hi - row_lo, p.seg_rows), 1);
  last_launch_shape() = {ctas_x, kNumSms * ctas_per_sm};
  SODA_LAUNCH(kernel, grid, kWarps * 32, S::kBytes, a.stream, p);
  launch_counter().fetch_add(1);
  SODA_CUDA_CHECK(cudaGetLastError());
  return SODA_CUDA_OK;
}

template <class Prog>
int launch_pass_2d(const PassArgs& a) {
  if constexpr (Prog::kWarps > 1) {
    int x_lo = a.extent[0], x_hi = 0;
    for (int o = 0; o < Prog::kNumOutputs; ++o) {
      if (a.box_lo[o][0] < x_lo) x_lo = a.box_lo[o][0];
      if (a.box_hi[o][0] > x_hi) x_hi = a.box_hi[o][0];
    }
    const int strips =
        x_hi > x_lo ? ceil_div(x_hi - floor_to(x_lo, Prog::kAlign0), Prog::kValid0)
                    : 0;
    if (strips <= narrow_grid_strips()) return launch_pass_2d_as<Prog, 1>(a);
  }
  return launch_pass_2d_as<Prog, Prog::kWarps>(a);
}

template <class Prog>
int launch_pass_3d(const PassArgs& a) {
  using S = Smem3D<Prog>;
  auto kernel = soda_stream3d_kernel<Prog>;
  int ctas_per_sm = 1;
  {
    int status = kernel_setup(kernel, Prog::kWarps * 32, S::kBytes, &ctas_per_sm);
    if (status != SODA_CUDA_OK) return status;
  }

  Params3D<Prog> p;
  memset(&p, 0, sizeof(p));
  const int box[3] = {Prog::kStrip, Prog::kRows, 1};
  int status = make_input_maps<Prog>(p.in_map, a, box);
  if (status != SODA_CUDA_OK) return status;

  int lo[3] = {a.extent[0], a.extent[1], a.extent[2]}, hi[3] = {0, 0, 0};
  for (int o = 0; o < Prog::kNumOutputs; ++o) {
    p.out[o] = a.out[o];
    p.out_pitch[o] = a.out_pitch[o][0];
    p.out_plane_pitch[o] = a.out_pitch[o][1];
    for (int d = 0; d < 3; ++d) {
      p.box_lo[o][d] = a.box_lo[o][d];
      p.box_hi[o][d] = a.box_hi[o][d];
      if (a.box_lo[o][d] < lo[d]) lo[d] = a.box_lo[o][d];
      if (a.box_hi[o][d] > hi[d]) hi[d] = a.box_hi[o][d];
    }
  }
  for (int d = 0; d < 3; ++d)
    if (hi[d] <= lo[d]) return SODA_CUDA_OK;

  p.x_origin = floor_to(lo[0], Prog::kAlign0);
  p.y_origin = lo[1];
  p.plane_lo = lo[2];
  p.plane_hi = hi[2];
  const int tiles_x = ceil_div(hi[0] - p.x_origin, Prog::kValid0);
  const int tiles_y = ceil_div(hi[1] - lo[1], Prog::kValid1);
  p.seg_planes = choose_segment(hi[2] - lo[2], tiles_x * tiles_y,
                                Prog::kMaxLag - Prog::kLoS, ctas_per_sm, a.segment);
  p.vec_ok = outputs_vector_aligned<Prog>(a) ? 1 : 0;
  dim3 grid(tiles_x, tiles_y, ceil_div(hi[2] - lo[2], p.seg_planes));
  last_launch_shape() = {tiles_x * tiles_y, kNumSms * ctas_per_sm};
  SODA_LAUNCH(kernel, grid, Prog::kWarps * 32, S::kBytes, a.stream, p);
  launch_counter().fetch_add(1);
  SODA_CUDA_CHECK(cudaGetLastError());
  return SODA_CUDA_OK;
}

template <class Prog>
int launch_pass(const PassArgs& a) {
  if constexpr (Prog::kDim == 2) {
    return launch_pass_2d<Prog>(a);
  } else {
    return launch_pass_3d<Prog>(a);
  }
}

template <class Prog>
soda_cuda_pass_info pass_info_of() {
  soda_cuda_pass_info info;
  memset(&info, 0, sizeof(info));
  info.time_block = Prog::kTimeBlock;
  for (int d = 0; d < Prog::kDim; ++d) {
    info.reach_lo[d] = Prog::kReachLo[d];
    info.reach_hi[d] = Prog::kReachHi[d];
  }
  info.cells_per_lane = Prog::kCells;
  info.strip_cells = Prog::kStrip;
  info.valid_cells[0] = Prog::kValid0;
  if constexpr (Prog::kDim == 2) {
    info.threads_per_cta = Prog::kWarps * 32;
    info.smem_bytes = Smem2D<Prog>::kBytes;
  } else {
    info.threads_per_cta = Prog::kWarps * 32;
    info.smem_bytes = Smem3D<Prog>::kBytes;
    info.valid_cells[1] = Prog::kValid1;
  }
  return info;
}

template <class Prog>
PassImpl make_pass_impl() {
  PassImpl impl;
  impl.info = pass_info_of<Prog>();
  impl.launch = &launch_pass<Prog>;
  impl.warmup = Prog::kMaxLag - Prog::kLoS;
  return impl;
}

// ---- measured segment length ---------------------------------------------------
// How many slices a CTA should stream depends on how the grid of CTAs fills
// the GPU (waves, tail), on the warm-up slices per segment and on whether the
// kernel is bound by HBM or by issue slots; the sweep in
// profiles/r01_segment_sweep.txt shows no simple rule.  So the first launch of
// a pass variant on a large grid times a handful of segment lengths with CUDA
// events on the caller's stream (the pass is idempotent: same inputs, same
// outputs) and every later launch of the same shape uses the fastest.
// opts->segment > 0 or SODA_CUDA_AUTOTUNE=0 turn this off.
struct TuneKey {
  int variant, device, extent[kMaxD], lo, hi;
  bool operator<(const TuneKey& o) const {
    if (variant != o.variant) return variant < o.variant;
    if (device != o.device) return device < o.device;
    for (int d = 0; d < kMaxD; ++d)
      if (extent[d] != o.extent[d]) return extent[d] < o.extent[d];
    if (lo != o.lo) return lo < o.lo;
    return hi < o.hi;
  }
};

// SODA_CUDA_SEGMENT=N (profiling aid): the segment length of every launch whose
// caller did not choose one; takes the measured choice's place.
inline int env_segment() {
  static const int value = [] {
    const char* env = getenv("SODA_CUDA_SEGMENT");
    return env != nullptr ? atoi(env) : 0;
  }();
  return value;
}

inline bool autotune_enabled() {
  static const bool enabled = [] {
    const char* env = getenv("SODA_CUDA_AUTOTUNE");
    return env == nullptr || env[0] != '0';
  }();
  return enabled;
}

// Launches over fewer cells than this use the rule of choose_segment().
// SODA_CUDA_TUNE_MIN_CELLS_LOG2 overrides the exponent (experiments).
inline long long tune_min_cells() {
  static const long long value = [] {
    const char* env = getenv("SODA_CUDA_TUNE_MIN_CELLS_LOG2");
    return 1LL << (env != nullptr ? atoi(env) : 22);
  }();
  return value;
}

inline int launch_tuned(const ProgramDesc& prog, int variant, PassArgs a) {
  const PassImpl& impl = prog.impls[variant];
  const int dim = prog.info.dim, s_dim = dim - 1;
  if (a.segment == 0 && env_segment() > 0) a.segment = env_segment();
  long long cells = 1;
  for (int d = 0; d < dim; ++d) cells *= a.extent[d];
  int lo = a.extent[s_dim], hi = 0;
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    if (a.box_lo[o][s_dim] < lo) lo = a.box_lo[o][s_dim];
    if (a.box_hi[o][s_dim] > hi) hi = a.box_hi[o][s_dim];
  }
  const int slices = hi - lo;
  const int min_segment = impl.warmup * 2 > 8 ? impl.warmup * 2 : 8;
  // 4 Mi cells and up: the chunk windows of the host pipeline (1/16 to 1/64 of
  // a grid) are measured too.  By the rule of choose_segment() a 16384 x 640
  // window of the bench kernel is 4 segments = 136 CTAs, under one wave: the 11
  // passes of such a chunk took 0.94 ms against 0.42 ms with the measured
  // choice, more than the chunk's upload (0.74 ms), so that 24 chunks and
  // more ran compute-bound (profiles/r02_pipeline_chunk_compute_old_new_threshold.jsonl)
  if (a.segment != 0 || !autotune_enabled() || cells < tune_min_cells() ||
      slices < 4 * min_segment)
    return impl.launch(a);

  static std::mutex mutex;
  static std::map<TuneKey, int> cache;
  TuneKey key;
  memset(&key, 0, sizeof(key));
  key.variant = variant;
  cudaGetDevice(&key.device);
  for (int d = 0; d < dim; ++d) key.extent[d] = a.extent[d];
  // store boxes that differ by a few slices (the shrinking windows of an
  // exchange group, multi_gpu.py; the chunk windows of the host pipeline,
  // whose heights differ by one) share one measurement
  key.extent[s_dim] = a.extent[s_dim] / 64;
  key.lo = 0;
  key.hi = slices / 64;
  std::lock_guard<std::mutex> lock(mutex);
  auto found = cache.find(key);
  if (found != cache.end()) {
    a.segment = found->second;
    return impl.launch(a);
  }

  const long long counted = launch_counter().load();
  int status = impl.launch(a);  // untimed: module load, attributes, cold caches
  if (status != SODA_CUDA_OK) return status;
  // destroyed on every exit path
  struct Events {
    cudaEvent_t start = nullptr, stop = nullptr;
    ~Events() {
      if (start) cudaEventDestroy(start);
      if (stop) cudaEventDestroy(stop);
    }
  } events;
  SODA_CUDA_CHECK(cudaEventCreate(&events.start));
  SODA_CUDA_CHECK(cudaEventCreate(&events.stop));
  cudaEvent_t start = events.start, stop = events.stop;
  auto measure = [&](int segment, float* ms) -> int {
    PassArgs trial = a;
    trial.segment = segment;
    *ms = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {  // the minimum of five: boxes are noisy
      SODA_CUDA_CHECK(cudaEventRecord(start, a.stream));
      int st = impl.launch(trial);
      if (st != SODA_CUDA_OK) return st;
      SODA_CUDA_CHECK(cudaEventRecord(stop, a.stream));
      SODA_CUDA_CHECK(cudaEventSynchronize(stop));
      float t = 0;
      SODA_CUDA_CHECK(cudaEventElapsedTime(&t, start, stop));
      if (t < *ms) *ms = t;
    }
    return SODA_CUDA_OK;
  };
  int best_segment = 0;  // 0: the rule of choose_segment()
  float best_ms = 0;
  status = measure(0, &best_ms);
  // candidates: segment counts 1, 2, 3, 4, 6, 9, 13, ... plus the counts that
  // fill the GPU's resident-CTA slots exactly k times (a grid of one full wave
  // can beat its neighbours by 20 %), tried from many short segments to few
  // long ones; stop once long segments have become clearly slower
  std::vector<int> counts;
  for (int c = 1; slices / c >= min_segment; c = c < 4 ? c + 1 : c + c / 2)
    counts.push_back(c);
  const LaunchShape shape = last_launch_shape();
  if (shape.tiles > 0 && shape.slots > 0) {
    const int waves[] = {1, 2, 3, 4, 6, 8, 12, 16};
    for (int k : waves) {
      const int c = k * shape.slots / shape.tiles;
      if (c >= 1 && slices / c >= min_segment) counts.push_back(c);
    }
  }
  std::sort(counts.begin(), counts.end());
  int previous = -1, worse = 0;
  int second_segment = 0;  // runner-up of the scan
  float second_ms = best_ms;
  for (size_t i = counts.size(); i-- > 0 && status == SODA_CUDA_OK;) {
    const int segment = ceil_div(slices, counts[i]);
    if (segment == previous) continue;
    previous = segment;
    float ms = 0;
    status = measure(segment, &ms);
    if (status != SODA_CUDA_OK) break;
    if (ms < best_ms) {
      second_ms = best_ms;
      second_segment = best_segment;
      best_ms = ms;
      best_segment = segment;
      worse = 0;
    } else {
      if (ms < second_ms || second_segment == best_segment) {
        second_ms = ms;
        second_segment = segment;
      }
      if (ms > 1.5f * best_ms && ++worse >= 2) break;
    }
  }
  // The scan's winner and runner-up are often within the noise of five short
  // launches (boxes picked different lengths for the same kernel, +-5 % on the
  // pass): measure the two again, alternating, and keep the better one.
  if (status == SODA_CUDA_OK && second_segment != best_segment &&
      second_ms < 1.05f * best_ms) {
    float a = 1e30f, b = 1e30f;
    for (int round = 0; round < 3 && status == SODA_CUDA_OK; ++round) {
      float ms = 0;
      status = measure(best_segment, &ms);
      if (ms < a) a = ms;
      if (status == SODA_CUDA_OK) status = measure(second_segment, &ms);
      if (ms < b) b = ms;
    }
    if (status == SODA_CUDA_OK && b < a) {
      best_segment = second_segment;
      best_ms = b;
    } else if (status == SODA_CUDA_OK) {
      best_ms = a;
    }
  }
  launch_counter().store(counted + 1);  // the tuning launches repeat one pass
  if (status != SODA_CUDA_OK) return status;
  cache[key] = best_segment;
  if (getenv("SODA_CUDA_VERBOSE"))
    fprintf(stderr, "soda_cuda: variant %d slices %d -> segment %d (%.3f ms)\n",
            variant, slices, best_segment, best_ms);
  return SODA_CUDA_OK;
}

}  // namespace rt
}  // namespace soda

// the generated file defines this before including the runtime
static const soda::rt::ProgramDesc& soda_program();

// ---- plan ---------------------------------------------------------------------
struct soda_cuda_plan {
  int extent[soda::rt::kMaxD];  // extent of the plan's arrays (a slab: local)
  int device;
  cudaStream_t stream;
  int segment;
  // device staging for host-array calls (a slab: its local arrays)
  void* d_in[soda::rt::kMaxT];
  void* d_out[soda::rt::kMaxT];
  // ping-pong scratch between passes, per output
  void* scratch[2][soda::rt::kMaxT];
  long long pitch[2];  // elements between rows / planes of every plan buffer
  int host_chunks;     // requested chunk count of the pipelined host path, 0 = auto
  cudaStream_t copy_in_stream;   // H2D of chunk k+1 overlaps compute of chunk k
  cudaStream_t copy_out_stream;  // ... and D2H of chunk k-1
  // Slices of the streamed dimension, in the plan's own coordinates, in which
  // the final output is valid.  A stand-alone plan covers the whole grid:
  // [final_lo, extent - final_hi).  A slab's arrays are a window of a larger
  // grid, so the range comes from the global grid (it may start below 0 or
  // end beyond the local extent: only the global border clips).
  int s_valid_lo[soda::rt::kMaxT];
  int s_valid_hi[soda::rt::kMaxT];
};

namespace soda {
namespace rt {

inline long long plan_cells(const soda_cuda_plan* plan, int dim) {
  return dim == 2 ? plan->pitch[0] * plan->extent[1]
                  : plan->pitch[1] * plan->extent[2];
}

inline int plan_alloc(soda_cuda_plan* plan, void** ptr, int elem_bytes) {
  const ProgramDesc& prog = soda_program();
  if (*ptr != nullptr) return SODA_CUDA_OK;
  SODA_CUDA_CHECK(cudaMalloc(ptr, plan_cells(plan, prog.info.dim) * elem_bytes));
  return SODA_CUDA_OK;
}

// Store boxes of one pass over the plan's arrays: everything for an
// intermediate pass, the program's final valid box for the last one.
inline void default_boxes(const ProgramDesc& prog, const soda_cuda_plan* plan,
                          bool final_pass, int (*lo)[kMaxD], int (*hi)[kMaxD]) {
  const int s_dim = prog.info.dim - 1;
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    for (int d = 0; d < prog.info.dim; ++d) {
      lo[o][d] = final_pass ? prog.info.final_lo[o][d] : 0;
      hi[o][d] = plan->extent[d] - (final_pass ? prog.info.final_hi[o][d] : 0);
    }
    if (final_pass) {
      lo[o][s_dim] = std::max(0, plan->s_valid_lo[o]);
      hi[o][s_dim] = std::min(plan->extent[s_dim], plan->s_valid_hi[o]);
    }
  }
}

// What all passes together depend on along the streamed dimension.
inline void total_reach(const ProgramDesc& prog, int* reach_lo, int* reach_hi) {
  const int s_dim = prog.info.dim - 1;
  *reach_lo = *reach_hi = 0;
  for (int pass = 0; pass < prog.info.num_passes; ++pass) {
    const soda_cuda_pass_info& info = prog.impls[prog.schedule[pass]].info;
    // one-sided windows: a pass never reaches less than its own slice
    *reach_lo += std::max(0, -info.reach_lo[s_dim]);
    *reach_hi += std::max(0, info.reach_hi[s_dim]);
  }
}

// All passes for the output slices [s_begin, s_end) of the streamed dimension,
// on the sub-array [s_begin - reach_lo, s_end + reach_hi) of the device buffers
// (reach = what all passes together depend on).  Inside that window the
// dependency cone of every kept cell is complete, so the result is identical
// to running the whole grid at once; slices near the window edges compute
// garbage that is never stored to the output.  Intermediates ping-pong through
// the plan's scratch arrays (same window).
inline int run_passes_window(soda_cuda_plan* plan, const void* const* d_in,
                             const long long (*in_pitch)[2], void* const* d_out,
                             const long long (*out_pitch)[2], int s_begin,
                             int s_end) {
  const ProgramDesc& prog = soda_program();
  const int dim = prog.info.dim, s_dim = dim - 1;
  const int num_passes = prog.info.num_passes;
  const int n_in = prog.info.num_inputs, n_out = prog.info.num_outputs;
  int reach_lo = 0, reach_hi = 0;
  total_reach(prog, &reach_lo, &reach_hi);
  const int total = plan->extent[s_dim];
  const int view_lo = s_begin - reach_lo > 0 ? s_begin - reach_lo : 0;
  const int view_hi = s_end + reach_hi < total ? s_end + reach_hi : total;
  auto offset = [&](const void* base, const long long* pitch, int elem_bytes) {
    const long long slice_pitch = dim == 2 ? pitch[0] : pitch[1];
    return static_cast<const char*>(base) +
           static_cast<long long>(view_lo) * slice_pitch * elem_bytes;
  };
  for (int pass = 0; pass < num_passes; ++pass) {
    const bool first = pass == 0, last = pass == num_passes - 1;
    PassArgs a;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < dim; ++d) a.extent[d] = plan->extent[d];
    a.extent[s_dim] = view_hi - view_lo;
    a.segment = plan->segment;
    a.stream = plan->stream;
    for (int i = 0; i < n_in; ++i) {
      // pass inputs after the first are the previous pass's outputs
      const int elem = first ? prog.in_elem_bytes[i] : prog.out_elem_bytes[i];
      if (first) {
        a.in[i] = offset(d_in[i], in_pitch[i], elem);
        a.in_pitch[i][0] = in_pitch[i][0];
        a.in_pitch[i][1] = in_pitch[i][1];
      } else {
        a.in[i] = offset(plan->scratch[(pass - 1) & 1][i], plan->pitch, elem);
        a.in_pitch[i][0] = plan->pitch[0];
        a.in_pitch[i][1] = plan->pitch[1];
      }
    }
    for (int o = 0; o < n_out; ++o) {
      if (last) {
        a.out[o] = const_cast<char*>(
            offset(d_out[o], out_pitch[o], prog.out_elem_bytes[o]));
        a.out_pitch[o][0] = out_pitch[o][0];
        a.out_pitch[o][1] = out_pitch[o][1];
      } else {
        int status = plan_alloc(plan, &plan->scratch[pass & 1][o],
                                prog.out_elem_bytes[o]);
        if (status != SODA_CUDA_OK) return status;
        a.out[o] = const_cast<char*>(offset(plan->scratch[pass & 1][o],
                                            plan->pitch, prog.out_elem_bytes[o]));
        a.out_pitch[o][0] = plan->pitch[0];
        a.out_pitch[o][1] = plan->pitch[1];
      }
    }
    default_boxes(prog, plan, last, a.box_lo, a.box_hi);
    for (int o = 0; o < n_out; ++o) {
      if (last) {
        int lo = a.box_lo[o][s_dim] > s_begin ? a.box_lo[o][s_dim] : s_begin;
        int hi = a.box_hi[o][s_dim] < s_end ? a.box_hi[o][s_dim] : s_end;
        a.box_lo[o][s_dim] = lo - view_lo;
        a.box_hi[o][s_dim] = (hi > lo ? hi : lo) - view_lo;
      } else {
        a.box_lo[o][s_dim] = 0;
        a.box_hi[o][s_dim] = view_hi - view_lo;
      }
    }
    int status = launch_tuned(prog, prog.schedule[pass], a);
    if (status != SODA_CUDA_OK) return status;
  }
  return SODA_CUDA_OK;
}

inline int run_passes(soda_cuda_plan* plan, const void* const* d_in,
                      const long long (*in_pitch)[2], void* const* d_out,
                      const long long (*out_pitch)[2]) {
  const ProgramDesc& prog = soda_program();
  return run_passes_window(plan, d_in, in_pitch, d_out, out_pitch, 0,
                           plan->extent[prog.info.dim - 1]);
}

inline int check_strides(const int32_t* stride, const int* extent, int dim,
                         const char* what) {
  if (stride == nullptr) return SODA_CUDA_OK;  // dense
  if (stride[0] != 1)
    return fail(SODA_CUDA_UNSUPPORTED,
                std::string(what) + ": stride[0] must be 1 (dimension 0 dense)");
  for (int d = 1; d < dim; ++d) {
    if (static_cast<long long>(stride[d]) <
        static_cast<long long>(stride[d - 1]) * extent[d - 1])
      return fail(SODA_CUDA_BAD_ARGUMENT,
                  std::string(what) + ": strides overlap");
  }
  return SODA_CUDA_OK;
}

// Copies the box [lo, hi) (plan coordinates) between a strided host array and
// a plan buffer.  Slice s of the plan is slice s + host_shift of the host
// array (0 for a stand-alone plan; a slab's arrays start elsewhere than the
// caller's).
inline int copy_box(const soda_cuda_plan* plan, cudaStream_t stream, int dim,
                    void* device,
                    void* host, const int32_t* host_stride, int elem_bytes,
                    const int* lo, const int* hi, bool to_device,
                    int host_shift = 0) {
  long long hs[kMaxD] = {1, plan->extent[0],
                         static_cast<long long>(plan->extent[0]) * plan->extent[1]};
  if (host_stride != nullptr)
    for (int d = 0; d < dim; ++d) hs[d] = host_stride[d];
  const size_t width = static_cast<size_t>(hi[0] - lo[0]) * elem_bytes;
  if (dim == 2) {
    char* d_ptr = static_cast<char*>(device) +
                  (plan->pitch[0] * lo[1] + lo[0]) * elem_bytes;
    char* h_ptr = static_cast<char*>(host) +
                  (hs[1] * (lo[1] + host_shift) + lo[0]) * elem_bytes;
    const size_t rows = hi[1] - lo[1];
    if (to_device) {
      SODA_CUDA_CHECK(cudaMemcpy2DAsync(d_ptr, plan->pitch[0] * elem_bytes, h_ptr,
                                        hs[1] * elem_bytes, width, rows,
                                        cudaMemcpyHostToDevice, stream));
    } else {
      SODA_CUDA_CHECK(cudaMemcpy2DAsync(h_ptr, hs[1] * elem_bytes, d_ptr,
                                        plan->pitch[0] * elem_bytes, width, rows,
                                        cudaMemcpyDeviceToHost, stream));
    }
    return SODA_CUDA_OK;
  }
  if (hs[2] % hs[1] != 0)
    return fail(SODA_CUDA_UNSUPPORTED,
                "3-D host arrays need stride[2] to be a multiple of stride[1]");
  cudaMemcpy3DParms parms;
  memset(&parms, 0, sizeof(parms));
  cudaPitchedPtr d_pitched = make_cudaPitchedPtr(
      device, plan->pitch[0] * elem_bytes, plan->pitch[0] * elem_bytes,
      plan->pitch[1] / plan->pitch[0]);
  cudaPitchedPtr h_pitched = make_cudaPitchedPtr(
      host, hs[1] * elem_bytes, hs[1] * elem_bytes, hs[2] / hs[1]);
  cudaPos d_pos = make_cudaPos(static_cast<size_t>(lo[0]) * elem_bytes, lo[1], lo[2]);
  cudaPos h_pos = make_cudaPos(static_cast<size_t>(lo[0]) * elem_bytes, lo[1],
                               lo[2] + host_shift);
  parms.srcPtr = to_device ? h_pitched : d_pitched;
  parms.dstPtr = to_device ? d_pitched : h_pitched;
  parms.srcPos = to_device ? h_pos : d_pos;
  parms.dstPos = to_device ? d_pos : h_pos;
  parms.extent = make_cudaExtent(width, hi[1] - lo[1], hi[2] - lo[2]);
  parms.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  SODA_CUDA_CHECK(cudaMemcpy3DAsync(&parms, stream));
  return SODA_CUDA_OK;
}

// ---- chunked host pipeline -------------------------------------------------------
// Host arrays in, host arrays out, through the plan's device arrays: the
// slices [own_lo, own_hi) of the plan are produced in chunks along the streamed
// dimension.  chunk k: H2D (copy-in stream) -> all passes on its window (plan
// stream) -> D2H of its valid interior (copy-out stream); the copies of
// neighbouring chunks overlap the compute.  Chunks overlap by the total reach
// of all passes, so the price is (reach / chunk) redundant compute.
//
// A stand-alone plan owns every slice and the host arrays hold all of them.
// A slab (soda_slab.cuh) owns the middle of its arrays: the host holds
// [host_lo, host_hi) (its own slices, or - one process driving several devices
// from one global host array - the ghost slices too), and ghost slices the
// host does not hold arrive from the neighbouring ranks: `after_edges` is
// called once the slices next to the slab boundaries are queued for upload
// (it starts the exchange, ordered after the copy-in stream), `before_edge`
// before the first chunk that reads ghost slices is computed (it makes the
// plan stream wait for the exchange).
struct HostPipeline {
  soda_cuda_plan* plan = nullptr;
  const void* const* in_ptrs = nullptr;
  const int32_t* const* in_strides = nullptr;
  void* const* out_ptrs = nullptr;
  const int32_t* const* out_strides = nullptr;
  int own_lo = 0, own_hi = 0;    // slices to produce (plan coordinates)
  int host_lo = 0, host_hi = 0;  // slices the host arrays hold
  int host_shift = 0;            // plan slice s is host slice s + host_shift
  bool ghosts_from_peers = false;
  // whether chunks that read ghost slices are computed last (a transport of
  // unknown speed) or in their natural place (NCCL over NVLink: the exchange
  // of a few slices is over long before the first chunk's own upload is)
  bool defer_edge_chunks = true;
  int (*after_edges)(void* user, cudaStream_t copy_in) = nullptr;
  int (*before_edge)(void* user, cudaStream_t compute) = nullptr;
  void* user = nullptr;
  // state between issue() and finish()
  std::vector<cudaEvent_t> events;
  int result = SODA_CUDA_OK;
  // SODA_CUDA_PIPELINE_TRACE=1: when every chunk's upload, compute and download
  // ended (ms after the call's first queued operation) on stderr, one JSON line
  // per call.  The pipeline's own events carry the times (created with timing
  // in this mode) plus one more record per download: a diagnostic, not for
  // timed runs.
  struct Trace {
    cudaEvent_t start = nullptr;
    std::vector<cudaEvent_t> marks[3];  // upload (piece order), compute, download (compute order)
    std::vector<int> order, rows;
  } trace;
  static bool trace_enabled() {
    static const bool enabled = [] {
      const char* env = getenv("SODA_CUDA_PIPELINE_TRACE");
      return env != nullptr && env[0] != '0';
    }();
    return enabled;
  }
  int mark(int which, cudaStream_t stream) {
    if (!trace_enabled()) return SODA_CUDA_OK;
    cudaEvent_t event;
    SODA_CUDA_CHECK(cudaEventCreate(&event));
    events.push_back(event);
    if (which < 0) {
      trace.start = event;
    } else {
      trace.marks[which].push_back(event);
    }
    SODA_CUDA_CHECK(cudaEventRecord(event, stream));
    return SODA_CUDA_OK;
  }
  void print_trace() {
    if (trace.start == nullptr) return;
    static const char* const names[3] = {"upload_end", "compute_end", "download_end"};
    std::string line = "{\"pipeline_trace\": {\"chunks\": " +
                       std::to_string(trace.order.size());
    auto list = [&](const char* name, const std::vector<int>& values) {
      line += std::string(", \"") + name + "\": [";
      for (size_t i = 0; i < values.size(); ++i)
        line += (i ? ", " : "") + std::to_string(values[i]);
      line += "]";
    };
    list("compute_order", trace.order);
    list("chunk_slices", trace.rows);
    for (int w = 0; w < 3; ++w) {
      line += std::string(", \"") + names[w] + "_ms\": [";
      for (size_t i = 0; i < trace.marks[w].size(); ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, trace.start, trace.marks[w][i]);
        char text[32];
        snprintf(text, sizeof(text), "%s%.3f", i ? ", " : "", ms);
        line += text;
      }
      line += "]";
    }
    fprintf(stderr, "%s}}\n", line.c_str());
  }

  int new_event(cudaEvent_t* event) {
    SODA_CUDA_CHECK(cudaEventCreateWithFlags(
        event, trace_enabled() ? cudaEventDefault : cudaEventDisableTiming));
    events.push_back(*event);
    return SODA_CUDA_OK;
  }

  // SODA_CUDA_CHUNK_WEIGHTS=w0,w1,... (experiments): that many chunks, of
  // lengths in these proportions, whatever the options ask for.
  static const std::vector<double>& env_weights() {
    static const std::vector<double> weights = [] {
      std::vector<double> w;
      const char* env = getenv("SODA_CUDA_CHUNK_WEIGHTS");
      while (env != nullptr && *env != 0) {
        char* end = nullptr;
        const double value = strtod(env, &end);
        if (end == env) break;
        if (value > 0) w.push_back(value);
        env = *end == ',' ? end + 1 : end;
      }
      return w;
    }();
    return weights;
  }

  // Automatic layout: every chunk is this much shorter than the one before it.
  // The call ends one chunk's passes and one chunk's download after the last
  // upload, so the last chunk should be short; a chunk may be at most about
  // d2h_time / h2d_time (0.92 on B200, both directions busy) of its
  // predecessor or the downloads queue up behind the uploads.
  // SODA_CUDA_CHUNK_DECAY overrides (1 = equal chunks).
  static double auto_decay() {
    static const double value = [] {
      const char* env = getenv("SODA_CUDA_CHUNK_DECAY");
      const double v = env != nullptr ? atof(env) : 0.92;
      return v > 0.5 && v <= 1.0 ? v : 1.0;
    }();
    return value;
  }

  // Whether this process has not yet run this chunk layout on this device and
  // grid (the segment tuner's cache is per process too: the one-shot entry
  // points make a new plan for every call and must not measure again).
  static bool first_call_with_layout(const soda_cuda_plan* plan, int chunks,
                                     int lo, int hi) {
    static std::mutex mutex;
    static std::map<std::vector<int>, bool> seen;
    int device = 0;
    cudaGetDevice(&device);
    const std::vector<int> key = {device, plan->extent[0], plan->extent[1],
                                  plan->extent[2], chunks, lo, hi,
                                  plan->host_chunks};
    std::lock_guard<std::mutex> lock(mutex);
    return seen.emplace(key, true).second;
  }

  static int choose_chunks(const ProgramDesc& prog, const soda_cuda_plan* plan,
                           int slices) {
    if (env_weights().size() > 1 && static_cast<int>(env_weights().size()) <= slices)
      return static_cast<int>(env_weights().size());
    int reach_lo = 0, reach_hi = 0;
    total_reach(prog, &reach_lo, &reach_hi);
    const int reach = reach_lo + reach_hi;
    long long bytes = 0;
    const long long slice_cells =
        prog.info.dim == 2 ? plan->pitch[0] : plan->pitch[1];
    for (int i = 0; i < prog.info.num_inputs; ++i)
      bytes += slice_cells * slices * prog.in_elem_bytes[i];
    int chunks = plan->host_chunks;
    if (chunks < 0) {
      chunks = -chunks;  // this many chunks, shrinking towards the end (see issue())
    } else if (chunks == 0) {
      chunks = 1;
      if (bytes >= (32LL << 20)) {
        // chunks of at least 8x the reach: at most 12.5 % redundant compute at
        // the seams.  More chunks shorten the last download but not the call:
        // 16 / 24 / 32 / 48 / 64 chunks of the 16384^2 x 64 workload take 23.6 /
        // 23.5 / 23.7 / 24.3 / 24.8 ms, the uploads run back to back at 43-47
        // GB/s and lose about 0.07 ms per chunk to the passes' HBM traffic
        // (profiles/r02_e2e_chunks_tuner_threshold.jsonl, r02_pipeline_trace_summary.txt)
        // Redundant compute at the seams is cheap while compute hides under
        // the copies: an HBM-bound pass moves the grid at ~5 TB/s, the link
        // moves it at ~50 GB/s each way, so 16 passes may double (chunks as
        // short as the reach) and still take 0.6 of the copy time.  heat3d
        // 512^3 x 32 (16 passes, reach 64 of 512 planes): 1 / 4 / 8 chunks 19.7 /
        // 13.4 / 12.7 ms; jacobi2d 16384^2 x 256 (43 passes): 1 / 4 / 8 chunks
        // 52.8 / 29.7 / 29.2 ms (profiles/r02_e2e_programs_chunks.jsonl).
        // SODA_CUDA_CHUNK_REACH overrides the factor (experiments).
        static const int env_factor = [] {
          const char* env = getenv("SODA_CUDA_CHUNK_REACH");
          return env != nullptr ? atoi(env) : 0;
        }();
        const int factor =
            env_factor > 0 ? env_factor
            : prog.info.num_passes <= 16 ? 1
            : prog.info.num_passes <= 32 ? 2 : 4;
        chunks = slices / (factor * (reach > 0 ? reach : 1));
        if (chunks > 16) chunks = 16;
        if (chunks < 1) chunks = 1;
        // the automatic layout shrinks the chunks towards the end (issue()):
        // a quarter more of them
        if (chunks >= 8 && auto_decay() < 1.0) chunks += chunks / 4;
      }
    }
    if (chunks > slices) chunks = slices;
    return chunks < 1 ? 1 : chunks;
  }

  // Queues everything; nothing here waits for the device (pinned host memory).
  int issue() {
    const ProgramDesc& prog = soda_program();
    const int dim = prog.info.dim, s_dim = dim - 1;
    const int n_in = prog.info.num_inputs, n_out = prog.info.num_outputs;
    long long pitches[kMaxT][2];
    for (int i = 0; i < kMaxT; ++i) {
      pitches[i][0] = plan->pitch[0];
      pitches[i][1] = plan->pitch[1];
    }
    const int slices = own_hi - own_lo;
    if (slices <= 0) return SODA_CUDA_OK;
    const int chunks = choose_chunks(prog, plan, slices);
    int reach_lo = 0, reach_hi = 0;
    total_reach(prog, &reach_lo, &reach_hi);

    if (plan->copy_in_stream == nullptr) {
      SODA_CUDA_CHECK(cudaStreamCreateWithFlags(&plan->copy_in_stream,
                                                cudaStreamNonBlocking));
      SODA_CUDA_CHECK(cudaStreamCreateWithFlags(&plan->copy_out_stream,
                                                cudaStreamNonBlocking));
    }
    auto upload = [&](int lo_s, int hi_s) -> int {
      if (hi_s <= lo_s) return SODA_CUDA_OK;
      int c_lo[kMaxD] = {0, 0, 0}, c_hi[kMaxD];
      for (int d = 0; d < dim; ++d) c_hi[d] = plan->extent[d];
      c_lo[s_dim] = lo_s;
      c_hi[s_dim] = hi_s;
      for (int i = 0; i < n_in; ++i) {
        int status = copy_box(plan, plan->copy_in_stream, dim, plan->d_in[i],
                              const_cast<void*>(in_ptrs[i]),
                              in_strides ? in_strides[i] : nullptr,
                              prog.in_elem_bytes[i], c_lo, c_hi, true, host_shift);
        if (status != SODA_CUDA_OK) return status;
      }
      return SODA_CUDA_OK;
    };

    // the copy-in stream must not run ahead of work already queued on the plan
    // stream that still reads the staging buffers (a previous call)
    cudaEvent_t start_event;
    int status = new_event(&start_event);
    if (status != SODA_CUDA_OK) return status;
    SODA_CUDA_CHECK(cudaEventRecord(start_event, plan->stream));
    SODA_CUDA_CHECK(cudaStreamWaitEvent(plan->copy_in_stream, start_event, 0));
    status = mark(-1, plan->copy_in_stream);
    if (status != SODA_CUDA_OK) return status;

    const bool peers = ghosts_from_peers && after_edges != nullptr;
    if (peers) {
      // what the neighbours need from this rank goes first (a few slices,
      // uploaded again with their chunk), then the exchange runs beside the
      // remaining uploads
      status = upload(own_lo, std::min(own_hi, own_lo + reach_hi));
      if (status == SODA_CUDA_OK)
        status = upload(std::max(own_lo, own_hi - reach_lo), own_hi);
      if (status == SODA_CUDA_OK) status = after_edges(user, plan->copy_in_stream);
      if (status != SODA_CUDA_OK) return status;
    }

    // chunk bounds.  Equal chunks (a positive count in opts), or - the automatic
    // layout and a negative count - lengths in geometric progression: the
    // copy-in stream never waits and the copy-out stream runs one chunk
    // behind it, so the call ends one chunk's passes and one chunk's download
    // after the last upload and that last chunk should be short.  Measured on
    // B200 for the 16384^2 x 64 workload (profiles/r02_e2e_chunk_decay.jsonl,
    // alternating processes on one box): 16 equal chunks 24.1-25.1 ms, 20 chunks
    // of ratio 0.92 23.8-24.5 ms; ratios 0.90 / 0.94 and 16 / 24 / 28 chunks are
    // no better.  Shorter chunks at *both* ends (built and measured earlier)
    // gain nothing: when the first download starts does not matter.
    std::vector<int> bound(chunks + 1), piece(chunks + 1);
    const bool decay = plan->host_chunks < 0 ||
                       (plan->host_chunks == 0 && chunks >= 8 && auto_decay() < 1.0);
    if (static_cast<int>(env_weights().size()) == chunks && chunks > 1) {
      double total = 0, run = 0;
      for (double w : env_weights()) total += w;
      bound[0] = own_lo;
      for (int k = 0; k < chunks; ++k) {
        run += env_weights()[k];
        bound[k + 1] = own_lo + static_cast<int>(slices * (run / total) + 0.5);
        if (bound[k + 1] <= bound[k]) bound[k + 1] = bound[k] + 1;
      }
      bound[chunks] = own_hi;
    } else if (decay && chunks > 1) {
      // in steps of 64 slices where the chunks are long enough (the segment
      // tuner keeps one measurement per 64 slices of window height); no chunk
      // shorter than the reach of all passes
      const double r = auto_decay() < 1.0 ? auto_decay() : 0.92;
      double total = 0, run = 0, w = 1;
      for (int k = 0; k < chunks; ++k, w *= r) total += w;
      const int step = slices >= 64 * 4 * chunks ? 64 : 1;
      int shortest = std::max(step, reach_lo + reach_hi);
      if (static_cast<long long>(shortest) * chunks > slices) shortest = slices / chunks;
      bound[0] = own_lo;
      w = 1;
      for (int k = 0; k < chunks; ++k, w *= r) {
        run += w;
        int b = own_lo + static_cast<int>(slices * (run / total) / step + 0.5) * step;
        b = std::max(b, bound[k] + shortest);
        // what is left must hold the remaining chunks
        b = std::min(b, own_hi - (chunks - 1 - k) * shortest);
        bound[k + 1] = b;
      }
      bound[chunks] = own_hi;
    } else {
      for (int k = 0; k <= chunks; ++k)
        bound[k] = own_lo + static_cast<int>(static_cast<long long>(slices) * k / chunks);
    }
    // upload pieces: piece k ends where the window of chunk k ends (its upper
    // bound plus the reach of all passes), so that chunk k waits for its own
    // piece only, not for the whole upload of chunk k + 1; stretched to what
    // the host holds at both ends
    // The first call measures the segment length of every chunk window before
    // anything else is queued (on whatever the plan's arrays hold: every cell
    // these launches store is stored again below).  Left to the chunks' own
    // first launches, the measurements would run beside the uploads of the
    // chunks behind them, which share the HBM with the timed launches.
    const long long first_window_cells =
        (dim == 2 ? plan->extent[0]
                  : static_cast<long long>(plan->extent[0]) * plan->extent[1]) *
        (bound[1] - bound[0] + reach_lo + reach_hi);
    if (autotune_enabled() && plan->segment == 0 &&
        first_window_cells >= tune_min_cells() &&
        first_call_with_layout(plan, chunks, own_lo, own_hi)) {
      const long long counted = launch_counter().load();
      for (int k = 0; k < chunks && status == SODA_CUDA_OK; ++k)
        status = run_passes_window(plan, plan->d_in, pitches, plan->d_out, pitches,
                                   bound[k], bound[k + 1]);
      if (status != SODA_CUDA_OK) return status;
      SODA_CUDA_CHECK(cudaStreamSynchronize(plan->stream));
      launch_counter().store(counted);  // tuning, not work
    }
    for (int k = 0; k <= chunks; ++k)
      piece[k] = std::min(host_hi, std::max(host_lo, bound[k] + reach_hi));
    piece[0] = host_lo;
    piece[chunks] = host_hi;
    std::vector<cudaEvent_t> copied(chunks), computed(chunks);
    for (int k = 0; k < chunks; ++k) {
      status = new_event(&copied[k]);
      if (status == SODA_CUDA_OK) status = new_event(&computed[k]);
      if (status != SODA_CUDA_OK) return status;
    }
    for (int k = 0; k < chunks; ++k) {
      status = upload(piece[k], piece[k + 1]);
      if (status != SODA_CUDA_OK) return status;
      SODA_CUDA_CHECK(cudaEventRecord(copied[k], plan->copy_in_stream));
      if (trace_enabled()) trace.marks[0].push_back(copied[k]);
    }

    // compute order.  A chunk that reads ghost slices has to wait for the
    // exchange; with a transport of unknown speed such chunks go last.  With
    // a fast one they stay in place: chunk 0 of a rank with a neighbour below
    // would otherwise be uploaded first and computed last - one more upload
    // before the first compute and two more downloads after the last one
    // (N = 2, 16384^2 per rank: 32.5 ms per step).
    std::vector<int> order;
    size_t interior = 0;
    if (peers && defer_edge_chunks) {
      for (int k = 0; k < chunks; ++k) {
        const bool edge = bound[k] - reach_lo < host_lo ||
                          bound[k + 1] + reach_hi > host_hi;
        if (!edge) order.push_back(k);
      }
      interior = order.size();
      for (int k = 0; k < chunks; ++k)
        if (std::find(order.begin(), order.end(), k) == order.end())
          order.push_back(k);
    } else {
      for (int k = 0; k < chunks; ++k) order.push_back(k);
      // the exchange is waited for before the first chunk (it reads the ghost
      // slices below the slab or none at all)
      interior = peers ? 0 : order.size();
    }

    int lo[kMaxT][kMaxD], hi[kMaxT][kMaxD];
    default_boxes(prog, plan, true, lo, hi);
    bool waited_for_peers = false;
    for (size_t n = 0; n < order.size(); ++n) {
      const int k = order[n];
      if (n >= interior && !waited_for_peers && before_edge != nullptr) {
        status = before_edge(user, plan->stream);
        if (status != SODA_CUDA_OK) return status;
        waited_for_peers = true;
      }
      // the window of chunk k ends at bound[k+1] + reach_hi: wait for the last
      // piece it touches (uploads are issued in order on one stream)
      int last_needed = k;
      while (last_needed + 1 < chunks &&
             piece[last_needed + 1] < bound[k + 1] + reach_hi)
        ++last_needed;
      SODA_CUDA_CHECK(cudaStreamWaitEvent(plan->stream, copied[last_needed], 0));
      trace.order.push_back(k);
      trace.rows.push_back(bound[k + 1] - bound[k]);
      status = run_passes_window(plan, plan->d_in, pitches, plan->d_out, pitches,
                                 bound[k], bound[k + 1]);
      if (status != SODA_CUDA_OK) return status;
      SODA_CUDA_CHECK(cudaEventRecord(computed[k], plan->stream));
      if (trace_enabled()) trace.marks[1].push_back(computed[k]);
      SODA_CUDA_CHECK(cudaStreamWaitEvent(plan->copy_out_stream, computed[k], 0));
      for (int o = 0; o < n_out; ++o) {
        int o_lo[kMaxD], o_hi[kMaxD];
        bool empty = false;
        for (int d = 0; d < dim; ++d) {
          o_lo[d] = lo[o][d];
          o_hi[d] = hi[o][d];
        }
        if (o_lo[s_dim] < bound[k]) o_lo[s_dim] = bound[k];
        if (o_hi[s_dim] > bound[k + 1]) o_hi[s_dim] = bound[k + 1];
        for (int d = 0; d < dim; ++d) empty = empty || o_hi[d] <= o_lo[d];
        if (empty) continue;
        status = copy_box(plan, plan->copy_out_stream, dim, plan->d_out[o],
                          out_ptrs[o], out_strides ? out_strides[o] : nullptr,
                          prog.out_elem_bytes[o], o_lo, o_hi, false, host_shift);
        if (status != SODA_CUDA_OK) return status;
      }
      status = mark(2, plan->copy_out_stream);
      if (status != SODA_CUDA_OK) return status;
    }
    return SODA_CUDA_OK;
  }

  // Waits for everything issue() queued and releases its events.
  int finish(int issue_status) {
    cudaError_t sync_in = cudaSuccess, sync_out = cudaSuccess;
    if (plan->copy_in_stream) sync_in = cudaStreamSynchronize(plan->copy_in_stream);
    cudaError_t sync_compute = cudaStreamSynchronize(plan->stream);
    if (plan->copy_out_stream) sync_out = cudaStreamSynchronize(plan->copy_out_stream);
    if (issue_status == SODA_CUDA_OK && sync_in == cudaSuccess &&
        sync_compute == cudaSuccess && sync_out == cudaSuccess)
      print_trace();
    for (cudaEvent_t event : events) cudaEventDestroy(event);
    events.clear();
    if (issue_status != SODA_CUDA_OK) return issue_status;
    SODA_CUDA_CHECK(sync_in);
    SODA_CUDA_CHECK(sync_compute);
    SODA_CUDA_CHECK(sync_out);
    return SODA_CUDA_OK;
  }
};

struct DeviceGuard {
  int previous = -1;
  bool switched = false;
  int enter(int device) {
    if (device < 0) return SODA_CUDA_OK;
    SODA_CUDA_CHECK(cudaGetDevice(&previous));
    if (previous != device) {
      SODA_CUDA_CHECK(cudaSetDevice(device));
      switched = true;
    }
    return SODA_CUDA_OK;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(previous);
  }
};

}  // namespace rt
}  // namespace soda

// ---- the C ABI ----------------------------------------------------------------
extern "C" {

int soda_cuda_info(soda_cuda_program_info* info) {
  if (info == nullptr)
    return soda::rt::fail(SODA_CUDA_BAD_ARGUMENT, "info is NULL");
  *info = soda_program().info;
  return SODA_CUDA_OK;
}

int soda_cuda_get_pass_info(int32_t pass_index, soda_cuda_pass_info* info) {
  const soda::rt::ProgramDesc& prog = soda_program();
  if (info == nullptr || pass_index < 0 || pass_index >= prog.info.num_passes)
    return soda::rt::fail(SODA_CUDA_BAD_ARGUMENT, "bad pass index");
  *info = prog.impls[prog.schedule[pass_index]].info;
  return SODA_CUDA_OK;
}

const char* soda_cuda_last_error(void) { return soda::rt::last_error().c_str(); }

int64_t soda_cuda_launch_count(void) { return soda::rt::launch_counter().load(); }

int soda_cuda_set_param(int32_t index, const void* values,
                        const soda_cuda_opts* opts) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (index < 0 || index >= prog.info.num_params)
    return fail(SODA_CUDA_BAD_ARGUMENT, "bad param index");
  if (values == nullptr) return fail(SODA_CUDA_BAD_ARGUMENT, "param values are NULL");
  DeviceGuard guard;
  int status = guard.enter(opts ? opts->device : -1);
  if (status != SODA_CUDA_OK) return status;
  cudaStream_t stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  // ordered after the launches already queued on the stream, visible to the
  // ones queued afterwards
  SODA_CUDA_CHECK(cudaMemcpyToSymbolAsync(prog.param_symbol[index], values,
                                          prog.param_bytes[index], 0,
                                          cudaMemcpyHostToDevice, stream));
  SODA_CUDA_CHECK(cudaStreamSynchronize(stream));
  return SODA_CUDA_OK;
}

int soda_cuda_plan_create(const int32_t* extent, const soda_cuda_opts* opts,
                          soda_cuda_plan** out) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (extent == nullptr || out == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "extent / plan is NULL");
  for (int d = 0; d < prog.info.dim; ++d)
    if (extent[d] <= 0) return fail(SODA_CUDA_BAD_ARGUMENT, "extent must be positive");
  int device_count = 0;
  SODA_CUDA_CHECK(cudaGetDeviceCount(&device_count));
  if (device_count == 0)
    return fail(SODA_CUDA_CUDA_ERROR, "no CUDA device: this library has no CPU path");
  soda_cuda_plan* plan = new soda_cuda_plan();
  memset(plan, 0, sizeof(*plan));
  for (int d = 0; d < prog.info.dim; ++d) plan->extent[d] = extent[d];
  plan->device = -1;
  if (opts != nullptr) {
    plan->device = opts->device;
    plan->stream = static_cast<cudaStream_t>(opts->stream);
    plan->segment = opts->segment;
    plan->host_chunks = opts->reserved[0];
  }
  if (plan->device < 0) {
    cudaError_t err = cudaGetDevice(&plan->device);
    if (err != cudaSuccess) {
      delete plan;
      return fail(SODA_CUDA_CUDA_ERROR, cudaGetErrorString(err));
    }
  }
  // rows padded to 128 bytes for the widest element so TMA strides and vector
  // stores are aligned for every tensor of the program
  const long long row_align = 128;
  plan->pitch[0] = (extent[0] + row_align - 1) / row_align * row_align;
  plan->pitch[1] = prog.info.dim == 3 ? plan->pitch[0] * extent[1] : 0;
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    const int s_dim = prog.info.dim - 1;
    plan->s_valid_lo[o] = prog.info.final_lo[o][s_dim];
    plan->s_valid_hi[o] = extent[s_dim] - prog.info.final_hi[o][s_dim];
  }
  *out = plan;
  return SODA_CUDA_OK;
}

int soda_cuda_plan_destroy(soda_cuda_plan* plan) {
  using namespace soda::rt;
  if (plan == nullptr) return SODA_CUDA_OK;
  DeviceGuard guard;
  guard.enter(plan->device);
  for (int i = 0; i < kMaxT; ++i) {
    if (plan->d_in[i]) cudaFree(plan->d_in[i]);
    if (plan->d_out[i]) cudaFree(plan->d_out[i]);
    if (plan->scratch[0][i]) cudaFree(plan->scratch[0][i]);
    if (plan->scratch[1][i]) cudaFree(plan->scratch[1][i]);
  }
  if (plan->copy_in_stream) cudaStreamDestroy(plan->copy_in_stream);
  if (plan->copy_out_stream) cudaStreamDestroy(plan->copy_out_stream);
  delete plan;
  return SODA_CUDA_OK;
}

int soda_cuda_plan_run_device(soda_cuda_plan* plan, const void* const* d_in,
                              const int64_t (*in_pitches)[2], void* const* d_out,
                              const int64_t (*out_pitches)[2]) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (plan == nullptr || d_in == nullptr || d_out == nullptr ||
      in_pitches == nullptr || out_pitches == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "NULL argument");
  DeviceGuard guard;
  int status = guard.enter(plan->device);
  if (status != SODA_CUDA_OK) return status;
  long long ip[kMaxT][2], op[kMaxT][2];
  for (int i = 0; i < prog.info.num_inputs; ++i) {
    if (d_in[i] == nullptr) return fail(SODA_CUDA_BAD_ARGUMENT, "NULL input");
    ip[i][0] = in_pitches[i][0];
    ip[i][1] = in_pitches[i][1];
  }
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    if (d_out[o] == nullptr) return fail(SODA_CUDA_BAD_ARGUMENT, "NULL output");
    op[o][0] = out_pitches[o][0];
    op[o][1] = out_pitches[o][1];
  }
  return run_passes(plan, d_in, ip, d_out, op);
}

// Checks the caller's arrays and allocates the plan's staging buffers.
static int soda_plan_prepare_host(soda_cuda_plan* plan, const void* const* in_ptrs,
                                  const int32_t* const* in_strides,
                                  void* const* out_ptrs,
                                  const int32_t* const* out_strides,
                                  const int* host_extent) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  const int dim = prog.info.dim;
  for (int i = 0; i < prog.info.num_inputs; ++i) {
    if (in_ptrs[i] == nullptr) return fail(SODA_CUDA_BAD_ARGUMENT, "NULL input");
    const int32_t* stride = in_strides ? in_strides[i] : nullptr;
    int status = check_strides(stride, host_extent, dim, prog.info.input_names[i]);
    if (status != SODA_CUDA_OK) return status;
    status = plan_alloc(plan, &plan->d_in[i], prog.in_elem_bytes[i]);
    if (status != SODA_CUDA_OK) return status;
  }
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    if (out_ptrs[o] == nullptr) return fail(SODA_CUDA_BAD_ARGUMENT, "NULL output");
    const int32_t* stride = out_strides ? out_strides[o] : nullptr;
    int status = check_strides(stride, host_extent, dim, prog.info.output_names[o]);
    if (status != SODA_CUDA_OK) return status;
    status = plan_alloc(plan, &plan->d_out[o], prog.out_elem_bytes[o]);
    if (status != SODA_CUDA_OK) return status;
  }
  return SODA_CUDA_OK;
}

int soda_cuda_plan_run_host(soda_cuda_plan* plan, const void* const* in_ptrs,
                            const int32_t* const* in_strides,
                            void* const* out_ptrs,
                            const int32_t* const* out_strides) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (plan == nullptr || in_ptrs == nullptr || out_ptrs == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "NULL argument");
  const int dim = prog.info.dim;
  DeviceGuard guard;
  int status = guard.enter(plan->device);
  if (status != SODA_CUDA_OK) return status;
  status = soda_plan_prepare_host(plan, in_ptrs, in_strides, out_ptrs, out_strides,
                                  plan->extent);
  if (status != SODA_CUDA_OK) return status;

  const int zero[kMaxD] = {0, 0, 0};
  long long pitches[kMaxT][2];
  for (int i = 0; i < kMaxT; ++i) {
    pitches[i][0] = plan->pitch[0];
    pitches[i][1] = plan->pitch[1];
  }
  const int total = plan->extent[dim - 1];
  if (HostPipeline::choose_chunks(prog, plan, total) == 1) {
    int lo[kMaxT][kMaxD], hi[kMaxT][kMaxD];
    default_boxes(prog, plan, true, lo, hi);
    for (int i = 0; i < prog.info.num_inputs; ++i) {
      status = copy_box(plan, plan->stream, dim, plan->d_in[i],
                        const_cast<void*>(in_ptrs[i]),
                        in_strides ? in_strides[i] : nullptr,
                        prog.in_elem_bytes[i], zero, plan->extent, true);
      if (status != SODA_CUDA_OK) return status;
    }
    status = run_passes(plan, plan->d_in, pitches, plan->d_out, pitches);
    if (status != SODA_CUDA_OK) return status;
    // only the valid interior goes back; the rest of the caller's array is
    // untouched
    for (int o = 0; o < prog.info.num_outputs; ++o) {
      bool empty = false;
      for (int d = 0; d < dim; ++d) empty = empty || hi[o][d] <= lo[o][d];
      if (empty) continue;
      status = copy_box(plan, plan->stream, dim, plan->d_out[o], out_ptrs[o],
                        out_strides ? out_strides[o] : nullptr,
                        prog.out_elem_bytes[o], lo[o], hi[o], false);
      if (status != SODA_CUDA_OK) return status;
    }
    SODA_CUDA_CHECK(cudaStreamSynchronize(plan->stream));
    return SODA_CUDA_OK;
  }

  HostPipeline pipe;
  pipe.plan = plan;
  pipe.in_ptrs = in_ptrs;
  pipe.in_strides = in_strides;
  pipe.out_ptrs = out_ptrs;
  pipe.out_strides = out_strides;
  pipe.own_lo = pipe.host_lo = 0;
  pipe.own_hi = pipe.host_hi = total;
  return pipe.finish(pipe.issue());
}

int soda_cuda_run_host(const void* const* in_ptrs, const int32_t* const* in_strides,
                       void* const* out_ptrs, const int32_t* const* out_strides,
                       const int32_t* extent, const soda_cuda_opts* opts) {
  if (opts != nullptr && opts->reserved[1] > 1)  // sodac --cuda-gpus N
    return soda_cuda_multi_run_host(in_ptrs, in_strides, out_ptrs, out_strides,
                                    extent, nullptr, opts->reserved[1], opts);
  soda_cuda_plan* plan = nullptr;
  int status = soda_cuda_plan_create(extent, opts, &plan);
  if (status != SODA_CUDA_OK) return status;
  status = soda_cuda_plan_run_host(plan, in_ptrs, in_strides, out_ptrs, out_strides);
  soda_cuda_plan_destroy(plan);
  return status;
}

int soda_cuda_run_pass(int32_t pass_index, const int32_t* extent,
                       const void* const* d_in, const int64_t (*in_pitches)[2],
                       void* const* d_out, const int64_t (*out_pitches)[2],
                       const int32_t (*box_lo)[SODA_CUDA_MAX_DIM],
                       const int32_t (*box_hi)[SODA_CUDA_MAX_DIM],
                       const soda_cuda_opts* opts) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (pass_index < 0 || pass_index >= prog.info.num_passes)
    return fail(SODA_CUDA_BAD_ARGUMENT, "bad pass index");
  if (extent == nullptr || d_in == nullptr || d_out == nullptr ||
      in_pitches == nullptr || out_pitches == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "NULL argument");
  DeviceGuard guard;
  int status = guard.enter(opts ? opts->device : -1);
  if (status != SODA_CUDA_OK) return status;
  PassArgs a;
  memset(&a, 0, sizeof(a));
  for (int d = 0; d < prog.info.dim; ++d) a.extent[d] = extent[d];
  a.segment = opts ? opts->segment : 0;
  a.stream = opts ? static_cast<cudaStream_t>(opts->stream) : nullptr;
  for (int i = 0; i < prog.info.num_inputs; ++i) {
    a.in[i] = d_in[i];
    a.in_pitch[i][0] = in_pitches[i][0];
    a.in_pitch[i][1] = in_pitches[i][1];
  }
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    a.out[o] = d_out[o];
    a.out_pitch[o][0] = out_pitches[o][0];
    a.out_pitch[o][1] = out_pitches[o][1];
  }
  if (box_lo != nullptr && box_hi != nullptr) {
    for (int o = 0; o < prog.info.num_outputs; ++o)
      for (int d = 0; d < prog.info.dim; ++d) {
        a.box_lo[o][d] = box_lo[o][d];
        a.box_hi[o][d] = box_hi[o][d];
      }
  } else {
    int ext[kMaxD] = {0, 0, 0};
    for (int d = 0; d < prog.info.dim; ++d) ext[d] = extent[d];
    soda_cuda_plan whole;
    memset(&whole, 0, sizeof(whole));
    for (int d = 0; d < prog.info.dim; ++d) whole.extent[d] = ext[d];
    for (int o = 0; o < prog.info.num_outputs; ++o) {
      whole.s_valid_lo[o] = prog.info.final_lo[o][prog.info.dim - 1];
      whole.s_valid_hi[o] =
          ext[prog.info.dim - 1] - prog.info.final_hi[o][prog.info.dim - 1];
    }
    default_boxes(prog, &whole, pass_index == prog.info.num_passes - 1, a.box_lo,
                  a.box_hi);
  }
  return launch_tuned(prog, prog.schedule[pass_index], a);
}

}  // extern "C"

#include "soda_slab.cuh"
