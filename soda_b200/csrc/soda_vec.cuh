// Register <-> memory moves of the kCells cells one lane owns.
#pragma once

namespace soda {

// ---- vectors ---------------------------------------------------------------
// kC consecutive cells of one tensor row, moved with one 4/8/16-byte access
// when the address allows it.
template <typename T, int kC>
struct alignas(sizeof(T) * kC >= 16 ? 16 : sizeof(T) * kC) Vec {
  T v[kC];
};

template <typename T, int kC>
__device__ __forceinline__ void load_shared_vec(T (&dst)[kC], const T* src) {
  constexpr int kBytes = sizeof(T) * kC;
  if constexpr (kBytes <= 16) {
    Vec<T, kC> tmp = *reinterpret_cast<const Vec<T, kC>*>(src);
#pragma unroll
    for (int i = 0; i < kC; ++i) dst[i] = tmp.v[i];
  } else {
    static_assert(kBytes % 16 == 0, "cells per lane must fill 16-byte vectors");
    constexpr int kPer = 16 / sizeof(T);
#pragma unroll
    for (int k = 0; k < kC / kPer; ++k) {
      Vec<T, kPer> tmp =
          *reinterpret_cast<const Vec<T, kPer>*>(src + k * kPer);
#pragma unroll
      for (int i = 0; i < kPer; ++i) dst[k * kPer + i] = tmp.v[i];
    }
  }
}

template <typename T, int kC>
__device__ __forceinline__ void store_global_vec(T* dst, const T (&src)[kC]) {
  constexpr int kBytes = sizeof(T) * kC;
  if constexpr (kBytes <= 16) {
    Vec<T, kC> tmp;
#pragma unroll
    for (int i = 0; i < kC; ++i) tmp.v[i] = src[i];
    *reinterpret_cast<Vec<T, kC>*>(dst) = tmp;
  } else {
    constexpr int kPer = 16 / sizeof(T);
#pragma unroll
    for (int k = 0; k < kC / kPer; ++k) {
      Vec<T, kPer> tmp;
#pragma unroll
      for (int i = 0; i < kPer; ++i) tmp.v[i] = src[k * kPer + i];
      *reinterpret_cast<Vec<T, kPer>*>(dst + k * kPer) = tmp;
    }
  }
}

}  // namespace soda
