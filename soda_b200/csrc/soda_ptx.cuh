// sm_100a primitives used by the SODA stencil templates: mbarrier, TMA tiled
// loads (cp.async.bulk.tensor), warp shuffles and vector global stores.
// Everything here is a thin inline-PTX wrapper; no algorithm lives in this file.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <type_traits>

// `param` arrays of a program (generated code) live in constant memory
#define SODA_CONSTANT __constant__

namespace soda {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ unsigned char* dyn_smem() {
  extern __shared__ __align__(1024) unsigned char soda_dyn_smem[];
  return soda_dyn_smem;
}

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}

// ---- mbarrier -----------------------------------------------------------
struct alignas(8) Mbarrier {
  uint64_t state;
};

__device__ __forceinline__ void mbar_init(Mbarrier* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
}

// makes mbarrier.init visible to the async proxy (TMA)
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// orders prior generic-proxy accesses to shared memory before later
// async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(Mbarrier* bar,
                                                      uint32_t bytes) {
  asm volatile(
      "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
          smem_u32(bar)),
      "r"(bytes)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(Mbarrier* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SODA_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SODA_MBAR_DONE;\n"
      "bra SODA_MBAR_WAIT;\n"
      "SODA_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA ----------------------------------------------------------------
using TensorMap = CUtensorMap;

__device__ __forceinline__ void tma_prefetch_desc(const TensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(
                   reinterpret_cast<uint64_t>(map))
               : "memory");
}

// Loads the box whose low corner is (c0, c1) into shared memory; out-of-range
// elements are zero-filled by the hardware.  Completion is signalled on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const TensorMap* map,
                                            int c0, int c1, Mbarrier* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::"
      "complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1),
      "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const TensorMap* map,
                                            int c0, int c1, int c2,
                                            Mbarrier* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::"
      "complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2),
      "r"(smem_u32(bar))
      : "memory");
}

// ---- warp shuffles ---------------------------------------------------------
// Value of `v` held by the lane `delta` positions to the right (delta > 0) or
// left (delta < 0).  Lanes whose source falls outside the warp get their own
// value back; the templates only ever store cells for which that cannot
// happen.
// Written as non-volatile inline PTX on purpose: the compiler may then merge
// shuffles of the same register by the same distance.  A 19-tap horizontal
// window asks for every neighbour cell once per tap and cell (152 shuffles per
// row of 8 cells in xcorr); after merging, every source cell is shuffled once
// (18).  The templates only shuffle in warp-uniform control flow.
__device__ __forceinline__ unsigned shfl_bits_down(unsigned v, int delta) {
  unsigned r;
  asm("shfl.sync.down.b32 %0, %1, %2, 0x1f, 0xffffffff;"
      : "=r"(r) : "r"(v), "r"(delta));
  return r;
}
__device__ __forceinline__ unsigned shfl_bits_up(unsigned v, int delta) {
  unsigned r;
  asm("shfl.sync.up.b32 %0, %1, %2, 0x0, 0xffffffff;"
      : "=r"(r) : "r"(v), "r"(delta));
  return r;
}

template <int kDelta, typename T>
__device__ __forceinline__ T shfl_rel(T v) {
  static_assert(kDelta != 0 && kDelta > -32 && kDelta < 32, "bad lane delta");
  static_assert(sizeof(T) <= 8, "shuffle of wide type");
  if constexpr (sizeof(T) == 8) {
    unsigned long long bits;
    memcpy(&bits, &v, 8);
    unsigned lo = static_cast<unsigned>(bits), hi = static_cast<unsigned>(bits >> 32);
    lo = kDelta > 0 ? shfl_bits_down(lo, kDelta) : shfl_bits_up(lo, -kDelta);
    hi = kDelta > 0 ? shfl_bits_down(hi, kDelta) : shfl_bits_up(hi, -kDelta);
    bits = (static_cast<unsigned long long>(hi) << 32) | lo;
    T r;
    memcpy(&r, &bits, 8);
    return r;
  } else if constexpr (sizeof(T) == 4) {
    unsigned bits;
    memcpy(&bits, &v, 4);
    bits = kDelta > 0 ? shfl_bits_down(bits, kDelta) : shfl_bits_up(bits, -kDelta);
    T r;
    memcpy(&r, &bits, 4);
    return r;
  } else if constexpr (!std::is_integral<T>::value) {
    // 16-bit floating cells (soda::half_t): the bit pattern travels
    unsigned short raw;
    memcpy(&raw, &v, sizeof(T));
    unsigned bits = raw;
    bits = kDelta > 0 ? shfl_bits_down(bits, kDelta) : shfl_bits_up(bits, -kDelta);
    raw = static_cast<unsigned short>(bits);
    T r;
    memcpy(&r, &raw, sizeof(T));
    return r;
  } else {
    // 8- and 16-bit cells travel sign- or zero-extended in a 32-bit register;
    // what arrives is again such a value (tell the compiler, it cannot see
    // through the shuffle and would re-normalise)
    const int w = static_cast<int>(v);
    const int r = static_cast<int>(
        kDelta > 0 ? shfl_bits_down(static_cast<unsigned>(w), kDelta)
                   : shfl_bits_up(static_cast<unsigned>(w), -kDelta));
    __builtin_assume(r == static_cast<int>(static_cast<T>(r)));
    return static_cast<T>(r);
  }
}

// ---- packed fp32 pairs (Blackwell FADD2 / FMUL2) ------------------------------
// Two IEEE fp32 values in one 64-bit register pair.  add/sub/mul round each
// half to nearest-even exactly like the scalar instructions, so results are
// bit-identical to scalar code; only the number of issue slots halves.
struct F2 {
  unsigned long long bits;
};

__device__ __forceinline__ F2 f2_pack(float lo, float hi) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.bits) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float f2_lo(F2 v) {
  float lo;
  asm("{ .reg .b32 hi; mov.b64 {%0, hi}, %1; }" : "=f"(lo) : "l"(v.bits));
  return lo;
}
__device__ __forceinline__ float f2_hi(F2 v) {
  float hi;
  asm("{ .reg .b32 lo; mov.b64 {lo, %0}, %1; }" : "=f"(hi) : "l"(v.bits));
  return hi;
}
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits));
  return r;
}
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) {
  F2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits));
  return r;
}
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever --fmad
// says (it does not do that to the scalar .rn forms), which would change
// results.  The product is therefore written as fma(a, b, -0.0) with the -0.0
// read from constant memory at run time: a * b + (-0.0) rounds to exactly the
// product (also for +-0 products), ptxas cannot fold an addend it does not
// know, and the FFMA2 cannot absorb a second add.
__constant__ unsigned long long soda_neg_zero_pair = 0x8000000080000000ull;

__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r.bits)
      : "l"(a.bits), "l"(b.bits), "l"(soda_neg_zero_pair));
  return r;
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r.bits)
      : "l"(a.bits), "l"(b.bits), "l"(c.bits));
  return r;
}
// fused multiply-add that --fmad=false does not touch (an explicit request)
__device__ __forceinline__ float fma_rn(float a, float b, float c) {
  return __fmaf_rn(a, b, c);
}
// A pair assembled from two registers that were loaded as part of a wider
// vector (LDS.128) is only a register copy to ptxas, which then re-assembles it
// with two MOVs at every use.  Adding a zero that is only known at run time
// makes each half a value in its own right, produced once, directly in place.
__constant__ int soda_zero_word = 0;

__device__ __forceinline__ F2 f2_pack_once(float lo, float hi) {
  const int z = soda_zero_word;
  return f2_pack(__int_as_float(__float_as_int(lo) + z),
                 __int_as_float(__float_as_int(hi) + z));
}

__device__ __forceinline__ F2 f2_neg(F2 a) {
  F2 r;
  r.bits = a.bits ^ 0x8000000080000000ull;
  return r;
}
template <int kDelta>
__device__ __forceinline__ F2 f2_shfl(F2 v) {
  F2 r;
  r.bits = shfl_rel<kDelta>(v.bits);
  return r;
}

// warp-uniform OR of a per-lane flag
__device__ __forceinline__ bool warp_any(bool flag) {
  return __any_sync(kFullMask, flag) != 0;
}

__device__ __forceinline__ void warp_sync() { __syncwarp(); }
__device__ __forceinline__ void cta_sync() { __syncthreads(); }
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

}  // namespace soda

#include "soda_vec.cuh"

// kernel launch (a macro so that the CPU emulation used by the tests can
// substitute its own launcher for the <<<>>> syntax)
#define SODA_LAUNCH(kernel, grid, threads, smem_bytes, stream, params) \
  kernel<<<grid, threads, smem_bytes, stream>>>(params)
