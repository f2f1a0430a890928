// TEST INFRASTRUCTURE ONLY - CPU emulation of the handful of CUDA primitives the
// SODA stencil templates use, so that the *same* template and runtime code
// (soda_stream.cuh, soda_runtime.cuh and the generated program) can be executed
// by the `-m "not gpu"` tests in a container without a GPU, under ASan/UBSan.
//
// This is not a CPU fallback: the product (soda_b200.codegen.cuda) never
// defines SODA_EMU, never includes this file and refuses to run without a
// CUDA device.  tests/emu/build_emu.py compiles a generated program with
//   g++ -std=c++20 -DSODA_EMU -I tests/emu ...
// Every CUDA thread becomes a std::thread; __syncthreads / warp shuffles are
// std::barrier rendezvous; TMA loads are synchronous box copies with zero fill
// that complete an emulated mbarrier.
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__

using std::max;
using std::min;

// ---- the slice of the CUDA runtime API used by soda_runtime.cuh ---------------
enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999 };
typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
// `param` arrays: plain host memory in the emulation
#define SODA_CONSTANT static
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
constexpr unsigned long long cudaEnableDefault = 0;

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// several emulated devices, so that the multi-device host path can be driven
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 4; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
// everything is synchronous in the emulation, so streams and events are inert
typedef void* cudaEvent_t;
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2;
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = reinterpret_cast<cudaStream_t>(0x1);
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = reinterpret_cast<cudaEvent_t>(0x1);
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaMemcpyToSymbolAsync(const void* symbol, const void* src,
                                           size_t bytes, size_t offset,
                                           cudaMemcpyKind, cudaStream_t) {
  memcpy(static_cast<char*>(const_cast<void*>(symbol)) + offset, src, bytes);
  return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
  return cudaEventCreateWithFlags(e, 0);
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  *ms = 1.0f;
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) {
  return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** ptr, size_t bytes) {
  // device buffers are deliberately filled with junk: the templates must never
  // let uninitialised scratch reach a stored cell
  size_t rounded = (bytes + 255) / 256 * 256;
  *ptr = aligned_alloc(256, rounded);
  if (*ptr == nullptr) return cudaErrorMemoryAllocation;
  memset(*ptr, 0xA5, rounded);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* ptr) { free(ptr); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* ptr, int value, size_t bytes, cudaStream_t) {
  memset(ptr, value, bytes);
  return cudaSuccess;
}
template <class K>
cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K>
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) {
  *n = 1;
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src,
                                     size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r)
    memcpy(static_cast<char*>(dst) + r * dpitch,
           static_cast<const char*>(src) + r * spitch, width);
  return cudaSuccess;
}
struct cudaPitchedPtr { void* ptr; size_t pitch, xsize, ysize; };
struct cudaPos { size_t x, y, z; };
struct cudaExtent { size_t width, height, depth; };
struct cudaMemcpy3DParms {
  cudaPitchedPtr srcPtr, dstPtr;
  cudaPos srcPos, dstPos;
  cudaExtent extent;
  cudaMemcpyKind kind;
};
inline cudaPitchedPtr make_cudaPitchedPtr(void* p, size_t pitch, size_t xs, size_t ys) {
  return cudaPitchedPtr{p, pitch, xs, ys};
}
inline cudaPos make_cudaPos(size_t x, size_t y, size_t z) { return cudaPos{x, y, z}; }
inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return cudaExtent{w, h, d}; }
inline cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms* p, cudaStream_t) {
  for (size_t z = 0; z < p->extent.depth; ++z)
    for (size_t y = 0; y < p->extent.height; ++y) {
      const char* s = static_cast<const char*>(p->srcPtr.ptr) +
                      ((p->srcPos.z + z) * p->srcPtr.ysize + p->srcPos.y + y) * p->srcPtr.pitch +
                      p->srcPos.x;
      char* d = static_cast<char*>(p->dstPtr.ptr) +
                ((p->dstPos.z + z) * p->dstPtr.ysize + p->dstPos.y + y) * p->dstPtr.pitch +
                p->dstPos.x;
      memcpy(d, s, p->extent.width);
    }
  return cudaSuccess;
}

// ---- the slice of the driver API: tensor maps -----------------------------------
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
enum CUresult { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
enum CUtensorMapDataType {
  CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_DATA_TYPE_UINT16,
  CU_TENSOR_MAP_DATA_TYPE_UINT32, CU_TENSOR_MAP_DATA_TYPE_INT32,
  CU_TENSOR_MAP_DATA_TYPE_UINT64, CU_TENSOR_MAP_DATA_TYPE_INT64,
  CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_DATA_TYPE_FLOAT64
};
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_256B };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE };

struct alignas(64) CUtensorMap {
  const char* base;
  int rank;
  int elem_bytes;
  int64_t extent[3];
  int64_t stride_bytes[3];
  int box[3];
};

inline CUresult soda_emu_encode_tiled(CUtensorMap* map, CUtensorMapDataType type,
                                      cuuint32_t rank, void* base,
                                      const cuuint64_t* dims, const cuuint64_t* strides,
                                      const cuuint32_t* box, const cuuint32_t* elem_strides,
                                      CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  static const int kBytes[] = {1, 2, 4, 4, 8, 8, 4, 8};
  map->base = static_cast<const char*>(base);
  map->rank = static_cast<int>(rank);
  map->elem_bytes = kBytes[type];
  // the constraints cuTensorMapEncodeTiled enforces
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return CUDA_ERROR_INVALID_VALUE;
  if ((box[0] * map->elem_bytes) % 16 != 0) return CUDA_ERROR_INVALID_VALUE;
  for (cuuint32_t d = 0; d < rank; ++d) {
    if (box[d] == 0 || box[d] > 256 || elem_strides[d] != 1) return CUDA_ERROR_INVALID_VALUE;
    if (dims[d] == 0) return CUDA_ERROR_INVALID_VALUE;
    map->extent[d] = static_cast<int64_t>(dims[d]);
    map->box[d] = static_cast<int>(box[d]);
    map->stride_bytes[d] = d == 0 ? map->elem_bytes : static_cast<int64_t>(strides[d - 1]);
    if (d > 0 && strides[d - 1] % 16 != 0) return CUDA_ERROR_INVALID_VALUE;
  }
  return CUDA_SUCCESS;
}

inline cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, unsigned long long,
                                           cudaDriverEntryPointQueryResult* result) {
  *fn = reinterpret_cast<void*>(&soda_emu_encode_tiled);
  *result = cudaDriverEntryPointSuccess;
  return cudaSuccess;
}

// ---- thread / block state ---------------------------------------------------------
struct SodaEmuCta {
  int threads;
  std::barrier<> cta_barrier;
  std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
  std::vector<std::vector<uint64_t>> warp_slots;  // [warp][lane]
  unsigned char* smem;
  explicit SodaEmuCta(int n) : threads(n), cta_barrier(n) {
    for (int w = 0; w < (n + 31) / 32; ++w) {
      int lanes = std::min(32, n - w * 32);
      warp_barrier.emplace_back(new std::barrier<>(lanes));
      warp_slots.emplace_back(32);
    }
  }
};

struct SodaEmuIdx { unsigned x, y, z; };
inline thread_local SodaEmuIdx threadIdx, blockIdx, blockDim, gridDim;
inline thread_local SodaEmuCta* soda_emu_cta = nullptr;

namespace soda {

constexpr unsigned kFullMask = 0xffffffffu;
using TensorMap = CUtensorMap;

inline unsigned char* dyn_smem() { return soda_emu_cta->smem; }

// ---- mbarrier -------------------------------------------------------------------------
struct alignas(8) Mbarrier {
  volatile uint32_t phases_done;
  int32_t pending_tx;
  int32_t pending_arrivals;
  int32_t count;
};
inline std::mutex& emu_mbar_mutex() {
  static std::mutex m;
  return m;
}
inline void emu_mbar_try_complete(Mbarrier* bar) {
  if (bar->pending_arrivals == 0 && bar->pending_tx == 0) {
    bar->pending_arrivals = bar->count;
    bar->phases_done = bar->phases_done + 1;
  }
}
inline void mbar_init(Mbarrier* bar, uint32_t count) {
  std::lock_guard<std::mutex> lock(emu_mbar_mutex());
  bar->phases_done = 0;
  bar->pending_tx = 0;
  bar->pending_arrivals = static_cast<int32_t>(count);
  bar->count = static_cast<int32_t>(count);
}
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
inline void mbar_arrive_expect_tx(Mbarrier* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> lock(emu_mbar_mutex());
  bar->pending_tx += static_cast<int32_t>(bytes);
  bar->pending_arrivals -= 1;
  emu_mbar_try_complete(bar);
}
inline void mbar_wait(Mbarrier* bar, uint32_t parity) {
  for (;;) {
    {
      std::lock_guard<std::mutex> lock(emu_mbar_mutex());
      if ((bar->phases_done & 1u) != parity) return;  // that phase completed
    }
    std::this_thread::yield();
  }
}

// ---- TMA ---------------------------------------------------------------------------------
inline void tma_prefetch_desc(const TensorMap*) {}

inline void emu_tma_load(void* dst, const TensorMap* map, const int* c, Mbarrier* bar) {
  char* out = static_cast<char*>(dst);
  const int eb = map->elem_bytes;
  int box[3] = {1, 1, 1};
  int64_t coord[3] = {0, 0, 0};
  for (int d = 0; d < map->rank; ++d) {
    box[d] = map->box[d];
    coord[d] = c[d];
  }
  // measured on B200: a box whose first element is not 16-byte aligned in
  // global memory raises "illegal instruction"
  if ((coord[0] * eb) % 16 != 0) {
    fprintf(stderr, "soda_emu: TMA box start %lld x %d bytes is not 16-byte aligned\n",
            static_cast<long long>(coord[0]), eb);
    abort();
  }
  for (int k = 0; k < box[2]; ++k)
    for (int j = 0; j < box[1]; ++j)
      for (int i = 0; i < box[0]; ++i) {
        const int64_t g[3] = {coord[0] + i, coord[1] + j, coord[2] + k};
        bool inside = true;
        int64_t offset = 0;
        for (int d = 0; d < map->rank; ++d) {
          inside = inside && g[d] >= 0 && g[d] < map->extent[d];
          offset += g[d] * map->stride_bytes[d];
        }
        char* cell = out + ((static_cast<int64_t>(k) * box[1] + j) * box[0] + i) * eb;
        if (inside) memcpy(cell, map->base + offset, eb);
        else memset(cell, 0, eb);
      }
  std::lock_guard<std::mutex> lock(emu_mbar_mutex());
  bar->pending_tx -= box[0] * box[1] * box[2] * eb;
  emu_mbar_try_complete(bar);
}
inline void tma_load_2d(void* dst, const TensorMap* map, int c0, int c1, Mbarrier* bar) {
  const int c[2] = {c0, c1};
  emu_tma_load(dst, map, c, bar);
}
inline void tma_load_3d(void* dst, const TensorMap* map, int c0, int c1, int c2,
                        Mbarrier* bar) {
  const int c[3] = {c0, c1, c2};
  emu_tma_load(dst, map, c, bar);
}

// ---- warp / CTA ----------------------------------------------------------------------------
inline int lane_id() { return threadIdx.x & 31; }
inline void warp_sync() { soda_emu_cta->warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
inline void cta_sync() { soda_emu_cta->cta_barrier.arrive_and_wait(); }

// packed fp32 pairs: two independent IEEE operations, as FADD2 / FMUL2 do
struct F2 {
  float lo, hi;
};
inline F2 f2_pack(float lo, float hi) { return F2{lo, hi}; }
inline F2 f2_pack_once(float lo, float hi) { return F2{lo, hi}; }
inline float f2_lo(F2 v) { return v.lo; }
inline float f2_hi(F2 v) { return v.hi; }
inline F2 f2_add(F2 a, F2 b) { return F2{a.lo + b.lo, a.hi + b.hi}; }
inline F2 f2_sub(F2 a, F2 b) { return F2{a.lo - b.lo, a.hi - b.hi}; }
inline F2 f2_mul(F2 a, F2 b) { return F2{a.lo * b.lo, a.hi * b.hi}; }
inline F2 f2_neg(F2 a) { return F2{-a.lo, -a.hi}; }
inline float fma_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline F2 f2_fma(F2 a, F2 b, F2 c) {
  return F2{std::fmaf(a.lo, b.lo, c.lo), std::fmaf(a.hi, b.hi, c.hi)};
}

inline bool warp_any(bool flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto& slots = soda_emu_cta->warp_slots[warp];
  slots[lane] = flag ? 1 : 0;
  warp_sync();
  bool any = false;
  const int lanes = std::min(32, soda_emu_cta->threads - warp * 32);
  for (int l = 0; l < lanes; ++l) any = any || slots[l] != 0;
  warp_sync();
  return any;
}

template <int kDelta, typename T>
inline T shfl_rel(T v) {
  static_assert(kDelta != 0 && kDelta > -32 && kDelta < 32, "bad lane delta");
  static_assert(sizeof(T) <= 8, "shuffle of wide type");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto& slots = soda_emu_cta->warp_slots[warp];
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  slots[lane] = raw;
  warp_sync();
  const int src = lane + kDelta;
  T result = v;  // out-of-range source: own value, like __shfl_up/down_sync
  if (src >= 0 && src < 32) memcpy(&result, &slots[src], sizeof(T));
  warp_sync();
  return result;
}

template <int kDelta>
inline F2 f2_shfl(F2 v) {
  return shfl_rel<kDelta>(v);
}

}  // namespace soda

#include "soda_vec.cuh"

// ---- launch ------------------------------------------------------------------------------------
template <class Kernel, class Params>
void soda_emu_launch(Kernel kernel, dim3 grid, int threads, size_t smem_bytes,
                     const Params& params) {
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        SodaEmuCta cta(threads);
        // guard bytes on both sides, poisoned, so ASan / junk values expose
        // any access outside the dynamic shared-memory window
        size_t rounded = (smem_bytes + 1023) / 1024 * 1024;
        unsigned char* raw = static_cast<unsigned char*>(aligned_alloc(1024, rounded + 1024));
        memset(raw, 0x5A, rounded + 1024);
        cta.smem = raw;
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) {
          pool.emplace_back([&, t] {
            threadIdx = SodaEmuIdx{static_cast<unsigned>(t), 0, 0};
            blockIdx = SodaEmuIdx{bx, by, bz};
            blockDim = SodaEmuIdx{static_cast<unsigned>(threads), 1, 1};
            gridDim = SodaEmuIdx{grid.x, grid.y, grid.z};
            soda_emu_cta = &cta;
            kernel(params);
          });
        }
        for (auto& th : pool) th.join();
        free(raw);
      }
}

#define SODA_LAUNCH(kernel, grid, threads, smem_bytes, stream, params) \
  soda_emu_launch(kernel, grid, threads, smem_bytes, params)
