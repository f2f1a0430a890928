// Register <-> memory moves of the kCells cells one lane owns.
#pragma once

namespace soda {

// ---- vectors ---------------------------------------------------------------
// kC consecutive cells of one tensor row, moved in pieces of kVecBytes: the
// largest power of two that divides the lane's bytes, at most 16 (24-byte
// lanes of 6 fp32 cells move as three 8-byte pieces).
__host__ __device__ constexpr int vec_piece_bytes(int bytes) {
  int piece = 1;
  while (piece < 16 && bytes % (piece * 2) == 0) piece *= 2;
  return piece;
}

template <typename T, int kC>
struct alignas(vec_piece_bytes(sizeof(T) * kC)) Vec {
  T v[kC];
};

template <typename T, int kC>
__device__ __forceinline__ void load_shared_vec(T (&dst)[kC], const T* src) {
  constexpr int kPiece = vec_piece_bytes(sizeof(T) * kC);
  constexpr int kPer = kPiece / int(sizeof(T)) > 0 ? kPiece / int(sizeof(T)) : 1;
  static_assert(kC % kPer == 0, "lane bytes must be whole pieces");
#pragma unroll
  for (int k = 0; k < kC / kPer; ++k) {
    Vec<T, kPer> tmp = *reinterpret_cast<const Vec<T, kPer>*>(src + k * kPer);
#pragma unroll
    for (int i = 0; i < kPer; ++i) dst[k * kPer + i] = tmp.v[i];
  }
}

template <typename T, int kC>
__device__ __forceinline__ void store_global_vec(T* dst, const T (&src)[kC]) {
  constexpr int kPiece = vec_piece_bytes(sizeof(T) * kC);
  constexpr int kPer = kPiece / int(sizeof(T)) > 0 ? kPiece / int(sizeof(T)) : 1;
#pragma unroll
  for (int k = 0; k < kC / kPer; ++k) {
    Vec<T, kPer> tmp;
#pragma unroll
    for (int i = 0; i < kPer; ++i) tmp.v[i] = src[k * kPer + i];
    *reinterpret_cast<Vec<T, kPer>*>(dst + k * kPer) = tmp;
  }
}

}  // namespace soda
