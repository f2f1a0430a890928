// Hand-written sm_100a stencil templates for SODA programs.
//
// A SODA program reaches these templates as a `Prog` type that carries only
//   * one functor per statement (the expression, printed from the IR), and
//   * constexpr plan tables (dependency DAG of one pass, lags, window sizes).
// Tiling, TMA pipelines, sliding windows, neighbour exchange, temporal
// blocking and stores are all written here, once.
//
// What the reference's FPGA micro-architecture becomes on a B200
// (reference: src/soda/core.py:684-795, src/soda/dataflow.py:555-603,
//  src/soda/codegen/xilinx/hls_kernel.py:665-886):
//   reuse / line buffers      -> per-thread register sliding window over the
//                                streamed (last) dimension: each lane keeps the
//                                last `ring` slices of the cells it owns
//   unrolled PEs fed by chains-> the 32 lanes of a warp own adjacent cells;
//                                dimension-0 neighbours arrive by warp shuffle
//   all stages + all iterate  -> every node of the pass DAG (time_block copies
//   copies chained spatially     of every statement) is evaluated per step, in
//                                registers, before anything returns to HBM
//   BurstRead / BurstWrite    -> TMA tiled loads into an mbarrier ring in
//                                shared memory; 16-byte coalesced stores
// A thread owns a patch of kCy rows x kCells cells (kCy == 1 in 2-D).
//
// Step t of a strip consumes input slice t; node n then produces its slice
// t - lag[n].  Values computed from cells outside the loaded tile are garbage
// by construction and are never stored: stores are restricted to the cells
// whose whole dependency cone was inside the tile (plan halos) and inside the
// grid (store box).
#pragma once

#include <type_traits>

#ifdef SODA_EMU
#include <soda_emu.h>  // CPU emulation of the primitives below (tests only)
#else
#include "soda_ptx.cuh"
#endif
#include "soda_half.cuh"

namespace soda {

constexpr int kMaxProds = 8;

struct NodeDesc {
  int kind;        // 0: pass input, 1: stage
  int src;         // input index, or functor index
  int lag;         // produces slice t - lag at step t
  int ring;        // slices kept in the register window (>= 1)
  int out;         // output array stored by this node, or -1
  int smem_depth;  // 3-D: slices readable by other warps (0: none)
  int smem_reach;  // 3-D: furthest dimension-1 offset at which it is read
  int nprod;
  int prod[kMaxProds];  // producer node per functor load slot
};

__host__ __device__ constexpr int floor_div(int a, int b) {
  return a >= 0 ? a / b : -((-a + b - 1) / b);
}

// ---- packed evaluation ----------------------------------------------------------
// Programs whose tensors are all fp32 and whose statements only add, subtract
// and multiply are evaluated two cells at a time (Prog::kPack == 2): every IR
// operation maps to one FADD2 / FFMA2 instead of two FADD / FMUL.  Rounding is
// per element and identical to the scalar instructions, so results do not
// change; the issue slots of the arithmetic halve.  The generated functors are
// the same text in both modes: they only use the operators and cast_to<>
// defined here.
//
// Pair layout ("half split"): with kCells = 2H cells per lane, unit u holds the
// lane's cells (u, u + H) - not two adjacent cells.  A load at dimension-0
// offset dx of unit u then is
//   * unit u + dx itself while 0 <= u + dx < H (no instruction at all), or
//   * a unit rotated across the lane boundary, (hi of unit j, lo of unit j of
//     the next lane): one 32-bit shuffle and one register move,
// whereas adjacent-cell pairs need two moves for every odd dx of every unit
// (measured on the jacobi2d time-block-5 kernel: 7.4 MOV per 10 packed
// arithmetic instructions before, 2 after).
__device__ __forceinline__ F2 f2_splat(float v) { return f2_pack(v, v); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return f2_add(a, b); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return f2_sub(a, b); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return f2_mul(a, b); }
__device__ __forceinline__ F2 operator-(F2 a) { return f2_neg(a); }
__device__ __forceinline__ F2 operator+(F2 a) { return a; }
// scalar operands are fp32 (or an int literal, which C++ converts to fp32 in a
// float expression): broadcast to both halves
template <typename S>
__device__ __forceinline__ F2 operator+(F2 a, S b) { return f2_add(a, f2_splat(float(b))); }
template <typename S>
__device__ __forceinline__ F2 operator+(S a, F2 b) { return f2_add(f2_splat(float(a)), b); }
template <typename S>
__device__ __forceinline__ F2 operator-(F2 a, S b) { return f2_sub(a, f2_splat(float(b))); }
template <typename S>
__device__ __forceinline__ F2 operator-(S a, F2 b) { return f2_sub(f2_splat(float(a)), b); }
template <typename S>
__device__ __forceinline__ F2 operator*(F2 a, S b) { return f2_mul(a, f2_splat(float(b))); }
template <typename S>
__device__ __forceinline__ F2 operator*(S a, F2 b) { return f2_mul(f2_splat(float(a)), b); }

// c * x + acc in one rounding, for a coefficient c that is a power of two
// (--cuda-pow2-fma).  The product c * x is then exact, so rounding c * x + acc
// once is what rounding the product and then the sum gives - except when
// c * x falls into the subnormal range (|x| < 2^-126 / c), where the separate
// product would lose bits: the result can then differ in the last place when
// acc is as tiny.  Opt-in for that reason; the functor emitter only uses it
// for float literals that are powers of two.
__device__ __forceinline__ float fma_pow2(float c, float x, float acc) {
  return fma_rn(c, x, acc);
}
__device__ __forceinline__ F2 fma_pow2(float c, F2 x, F2 acc) {
  return f2_fma(f2_splat(c), x, acc);
}
__device__ __forceinline__ F2 fma_pow2(float c, F2 x, float acc) {
  return f2_fma(f2_splat(c), x, f2_splat(acc));
}
__device__ __forceinline__ F2 fma_pow2(float c, float x, F2 acc) {
  return f2_fma(f2_splat(c), f2_splat(x), acc);
}

template <typename To, typename From>
__device__ __forceinline__ auto cast_to(From v) {
  if constexpr (std::is_same<From, F2>::value) {
    static_assert(std::is_same<To, float>::value, "packed values are fp32");
    return v;
  } else if constexpr (std::is_same<From, H2>::value) {
    static_assert(std::is_same<To, half_t>::value, "packed values are half");
    return v;
  } else {
    return To(v);
  }
}

// 8- and 16-bit integer cells are kept in 32-bit registers that hold the
// converted value (zero- or sign-extended), so that the conversion to the
// tensor's type is paid once, where the cell is produced, instead of as a mask
// or sign extension at every use (ptxas keeps 16-bit values in 32-bit
// registers and re-normalises them whenever it cannot see the upper bits).
template <typename T>
using CarrierOf = typename std::conditional<
    std::is_integral<T>::value && sizeof(T) < 4,
    typename std::conditional<std::is_signed<T>::value, int, unsigned>::type,
    T>::type;

// The pair type of a packed program's cells: fp32 pairs (F2, one 64-bit
// register pair, FADD2 / FFMA2) or binary16 pairs (H2, one 32-bit register,
// HADD2 / HMUL2).
template <typename T> struct PairOf { using type = T; };
template <> struct PairOf<float> { using type = F2; };
template <> struct PairOf<half_t> { using type = H2; };
__device__ __forceinline__ F2 pair_pack(float lo, float hi) { return f2_pack(lo, hi); }
__device__ __forceinline__ H2 pair_pack(half_t lo, half_t hi) { return h2_pack(lo, hi); }
// a pair that is a value in its own right (see f2_pack_once); a binary16 pair
// is one register anyway
__device__ __forceinline__ F2 pair_pack_once(float lo, float hi) { return f2_pack_once(lo, hi); }
__device__ __forceinline__ H2 pair_pack_once(half_t lo, half_t hi) { return h2_pack(lo, hi); }
__device__ __forceinline__ float pair_lo(F2 v) { return f2_lo(v); }
__device__ __forceinline__ float pair_hi(F2 v) { return f2_hi(v); }
__device__ __forceinline__ half_t pair_lo(H2 v) { return h2_lo(v); }
__device__ __forceinline__ half_t pair_hi(H2 v) { return h2_hi(v); }
template <int kDelta>
__device__ __forceinline__ F2 pair_shfl(F2 v) { return f2_shfl<kDelta>(v); }
template <int kDelta>
__device__ __forceinline__ H2 pair_shfl(H2 v) { return shfl_rel<kDelta>(v); }

// A lane owns kCells cells = kUnits units of kPack cells.
template <class Prog, int N>
using UnitOf = typename std::conditional<
    Prog::kPack == 2, typename PairOf<typename Prog::template T<N>>::type,
    CarrierOf<typename Prog::template T<N>>>::type;
template <class Prog>
constexpr int kUnitsOf = Prog::kCells / Prog::kPack;

template <class Prog, int N>
__device__ __forceinline__ void units_from_cells(
    UnitOf<Prog, N> (&units)[kUnitsOf<Prog>],
    const typename Prog::template T<N> (&cells)[Prog::kCells]) {
#pragma unroll
  for (int u = 0; u < kUnitsOf<Prog>; ++u) {
    if constexpr (Prog::kPack == 2) {
      units[u] = pair_pack_once(cells[u], cells[u + kUnitsOf<Prog>]);
    } else {
      units[u] = cells[u];
    }
  }
}

template <class Prog, int N>
__device__ __forceinline__ void cells_from_units(
    typename Prog::template T<N> (&cells)[Prog::kCells],
    const UnitOf<Prog, N> (&units)[kUnitsOf<Prog>]) {
#pragma unroll
  for (int u = 0; u < kUnitsOf<Prog>; ++u) {
    if constexpr (Prog::kPack == 2) {
      cells[u] = pair_lo(units[u]);
      cells[u + kUnitsOf<Prog>] = pair_hi(units[u]);
    } else {
      cells[u] = static_cast<typename Prog::template T<N>>(units[u]);
    }
  }
}

// ---- register sliding windows ------------------------------------------------
template <class Prog, int N>
struct RingStore : RingStore<Prog, N - 1> {
  UnitOf<Prog, N - 1> r[Prog::kNodes[N - 1].ring][Prog::kCy][kUnitsOf<Prog>];
};
template <class Prog>
struct RingStore<Prog, 0> {};

template <class Prog>
using Rings = RingStore<Prog, Prog::kNumNodes>;

template <int N, class Prog>
__device__ __forceinline__ auto& ring_of(Rings<Prog>& rings) {
  return static_cast<RingStore<Prog, N + 1>&>(rings).r;
}
template <int N, class Prog>
__device__ __forceinline__ const auto& ring_of(const Rings<Prog>& rings) {
  return static_cast<const RingStore<Prog, N + 1>&>(rings).r;
}

template <class Prog, int N = 0>
__device__ __forceinline__ void clear_rings(Rings<Prog>& rings) {
  if constexpr (N < Prog::kNumNodes) {
    using T = typename Prog::template T<N>;
    auto& r = ring_of<N, Prog>(rings);
#pragma unroll
    for (int s = 0; s < Prog::kNodes[N].ring; ++s) {
#pragma unroll
      for (int j = 0; j < Prog::kCy; ++j) {
#pragma unroll
        for (int u = 0; u < kUnitsOf<Prog>; ++u) {
          if constexpr (Prog::kPack == 2) {
            r[s][j][u] = pair_pack(T(0), T(0));
          } else {
            r[s][j][u] = T(0);
          }
        }
      }
    }
    clear_rings<Prog, N + 1>(rings);
  }
}

// Register slot that holds logical slot `logical` (0: oldest, ring-1: the slice
// of this step) of node P.  Contexts that know their phase within an unrolled
// loop (Ctx::kRotate) use the window as a circular buffer with compile-time
// indices whenever the unroll factor is a multiple of the depth: nothing is
// ever copied.  Otherwise the window is shifted every step.
template <class Prog, class Ctx, int P>
__host__ __device__ constexpr bool ring_rotates() {
  return Ctx::kRotate && Prog::kUnroll % Prog::kNodes[P].ring == 0;
}
template <class Prog, class Ctx, int P>
__host__ __device__ constexpr int phys_slot(int logical) {
  constexpr int kDepth = Prog::kNodes[P].ring;
  if (ring_rotates<Prog, Ctx, P>()) {
    return ((Ctx::kPhase - (kDepth - 1 - logical)) % kDepth + kDepth) % kDepth;
  }
  return logical;
}

// Oldest slice falls out, slot ring-1 becomes free for the slice of this step.
template <int N, class Prog, class Ctx>
__device__ __forceinline__ void advance_ring(Ctx& ctx) {
  if constexpr (!ring_rotates<Prog, Ctx, N>()) {
    auto& r = ring_of<N, Prog>(ctx.rings);
#pragma unroll
    for (int s = 0; s + 1 < Prog::kNodes[N].ring; ++s) {
#pragma unroll
      for (int j = 0; j < Prog::kCy; ++j) {
#pragma unroll
        for (int u = 0; u < kUnitsOf<Prog>; ++u) r[s][j][u] = r[s + 1][j][u];
      }
    }
  }
}

// The units of patch row J of the slice node N produces in this step.
template <int N, int J, class Prog, class Ctx>
__device__ __forceinline__ auto& newest_units(Ctx& ctx) {
  return ring_of<N, Prog>(ctx.rings)[phys_slot<Prog, Ctx, N>(
      Prog::kNodes[N].ring - 1)][J];
}

// ... as plain cells (for stores).
template <int N, int J, class Prog, class Ctx>
__device__ __forceinline__ void newest_cells(
    Ctx& ctx, typename Prog::template T<N> (&cells)[Prog::kCells]) {
  cells_from_units<Prog, N>(cells, newest_units<N, J, Prog>(ctx));
}

// ---- accessor handed to the generated functors --------------------------------
// ld<K, DX, DY, DS>() is the value of the K-th loaded tensor of the statement
// at offset (DX, DY, DS) from the unit being produced (DY is always 0 in 2-D;
// DS is the offset in the streamed dimension).  J is the row of the thread's
// patch, I the unit index within that row.  Cells the lane does not own come
// from the neighbouring lanes by warp shuffle; rows outside the patch (3-D)
// from shared-memory planes.
template <class Prog, class Ctx, int N, int J, int I>
struct Access {
  const Ctx& ctx;

  // one cell of producer P's slice in register slot kSlot, patch row kRow, by
  // cell index relative to the lane's first cell
  template <int P, int kSlot, int kRow, int kCell>
  __device__ __forceinline__ typename Prog::template T<P> cell() const {
    constexpr int kC = Prog::kCells;
    constexpr int kLane = floor_div(kCell, kC);
    constexpr int kLocal = kCell - kLane * kC;
    typename Prog::template T<P> v;
    if constexpr (Prog::kPack == 2) {
      constexpr int kH = kUnitsOf<Prog>;
      const auto unit = ring_of<P, Prog>(ctx.rings)[kSlot][kRow][kLocal % kH];
      v = kLocal >= kH ? pair_hi(unit) : pair_lo(unit);
    } else {
      v = static_cast<typename Prog::template T<P>>(
          ring_of<P, Prog>(ctx.rings)[kSlot][kRow][kLocal]);
    }
    if constexpr (kLane == 0) {
      return v;
    } else {
      return shfl_rel<kLane>(v);
    }
  }

  template <int K, int DX, int DY, int DS>
  __device__ __forceinline__ auto ld() const {
    static_assert(K < Prog::kNodes[N].nprod, "functor loads an undeclared slot");
    constexpr int P = Prog::kNodes[N].prod[K];
    constexpr int kDistance = Prog::kNodes[N].lag - Prog::kNodes[P].lag - DS;
    static_assert(kDistance >= Prog::kSkew,
                  "plan: consumer runs ahead of its producer");
    // first cell of the unit, relative to the lane's first cell (packed: the
    // second one is kHalf cells further)
    constexpr int kFirst = I + DX;
    constexpr int kHalf = kUnitsOf<Prog>;
    constexpr int kRow = J + DY;
    if constexpr (kRow >= 0 && kRow < Prog::kCy) {
      // pipelined plans (kSkew == 1) evaluate a node before its producers
      // advance in the same step: their newest slice is one step old
      constexpr int kLogical =
          Prog::kNodes[P].ring - 1 - (kDistance - Prog::kSkew);
      static_assert(kLogical >= 0, "plan: register window too short");
      constexpr int kSlot = phys_slot<Prog, Ctx, P>(kLogical);
      if constexpr (Prog::kPack == 1) {
        return cell<P, kSlot, kRow, kFirst>();
      } else {
        constexpr int kLane = floor_div(kFirst, Prog::kCells);
        constexpr int kLocal = kFirst - kLane * Prog::kCells;
        if constexpr (kLocal < kHalf) {
          // a whole unit of this or a neighbouring lane
          const auto v = ring_of<P, Prog>(ctx.rings)[kSlot][kRow][kLocal];
          if constexpr (kLane == 0) {
            return v;
          } else {
            return pair_shfl<kLane>(v);
          }
        } else {
          // rotated across a lane boundary: (hi of unit j in lane kLane,
          // lo of unit j in lane kLane + 1)
          const auto v = ring_of<P, Prog>(ctx.rings)[kSlot][kRow][kLocal - kHalf];
          auto first = pair_hi(v);
          auto second = pair_lo(v);
          if constexpr (kLane != 0) first = shfl_rel<kLane>(first);
          if constexpr (kLane + 1 != 0) second = shfl_rel<kLane + 1>(second);
          return pair_pack(first, second);
        }
      }
    } else {
      static_assert(Prog::kDim == 3, "dimension-1 offsets need a 3-D program");
      static_assert(kDistance < Prog::kNodes[P].smem_depth,
                    "plan: shared-memory window too short");
      if constexpr (Prog::kPack == 1) {
        return ctx.template plane_cell<P, kDistance, kFirst, kRow>();
      } else {
        return ctx.template plane_unit<P, kDistance, kFirst, kRow>();
      }
    }
  }
};

// every unit of every patch row of node N
template <class Prog, class Ctx, int N, int K = 0>
__device__ __forceinline__ void eval_units(Ctx& ctx) {
  if constexpr (K < Prog::kCy * kUnitsOf<Prog>) {
    using T = typename Prog::template T<N>;
    using F = typename Prog::template StageF<Prog::kNodes[N].src>;
    constexpr int J = K / kUnitsOf<Prog>;
    constexpr int I = K % kUnitsOf<Prog>;
    newest_units<N, J, Prog>(ctx)[I] =
        cast_to<T>(F::eval(Access<Prog, Ctx, N, J, I>{ctx}));
    eval_units<Prog, Ctx, N, K + 1>(ctx);
  }
}

// Per-lane store plan of one output, fixed for the whole strip / tile: which
// cells of the lane's vector may be written, where slice 0 of the lane's
// vector lives and how far consecutive slices are apart.  Computed once; the
// per-step work is then one pointer bump, one slice-range test and one
// (predicated) vector store.
struct StorePlan {
  unsigned char* ptr;   // address of the lane's vector in the current slice
  long long step;       // bytes between consecutive slices
  int slice_lo;         // slices [slice_lo, slice_hi) are stored
  int slice_hi;
  int mode;             // 0: nothing, 1: whole vector, 2: some cells (mask)
  unsigned mask;        // cells to store when mode == 2
};

template <typename T, int kC>
__device__ __forceinline__ void init_store_plan(StorePlan& plan, void* base,
                                                long long offset_elems,
                                                long long step_elems, int col,
                                                bool lane_ok, int box_lo,
                                                int box_hi, bool vec_ok,
                                                int slice_lo, int slice_hi,
                                                int first_slice) {
  plan.step = step_elems * static_cast<long long>(sizeof(T));
  plan.ptr = static_cast<unsigned char*>(base) +
             (offset_elems + col + first_slice * step_elems) *
                 static_cast<long long>(sizeof(T));
  plan.slice_lo = slice_lo;
  plan.slice_hi = slice_hi;
  plan.mask = 0;
#pragma unroll
  for (int i = 0; i < kC; ++i) {
    if (lane_ok && col + i >= box_lo && col + i < box_hi) plan.mask |= 1u << i;
  }
  if (plan.mask == 0) {
    plan.mode = 0;
  } else if (vec_ok && plan.mask == (1u << kC) - 1u) {
    plan.mode = 1;
  } else {
    plan.mode = 2;
  }
}

// Stores `v` as slice `slice` of the output and advances to the next slice.
// `any_partial` is warp-uniform: only warps that touch an edge of the store box
// take the per-cell path, everybody else issues a single predicated vector
// store and the warp stays converged for the shuffles that follow.
template <typename T, int kC>
__device__ __forceinline__ void store_slice(StorePlan& plan, const T (&v)[kC],
                                            int slice, bool any_partial) {
  T* dst = reinterpret_cast<T*>(plan.ptr);
  plan.ptr += plan.step;
  if (slice < plan.slice_lo || slice >= plan.slice_hi) return;  // uniform
  if (!any_partial) {
    if (plan.mode == 1) store_global_vec<T, kC>(dst, v);
  } else {
#pragma unroll
    for (int i = 0; i < kC; ++i) {
      if ((plan.mask >> i) & 1u) dst[i] = v[i];
    }
  }
}

// =============================================================================
// 2-D: every warp streams an independent column strip
// =============================================================================

template <class Prog>
struct Params2D {
  TensorMap in_map[Prog::kNumInputs];
  void* out[Prog::kNumOutputs];
  long long out_pitch[Prog::kNumOutputs];  // elements between rows
  // cells [box_lo, box_hi) of each output may be written (dim 0, dim 1)
  int box_lo[Prog::kNumOutputs][2];
  int box_hi[Prog::kNumOutputs][2];
  int x_origin;    // dimension-0 cell of strip 0's first valid cell
  int num_strips;
  int row_lo;      // first output row produced by segment 0
  int row_hi;      // one past the last output row produced
  int seg_rows;    // output rows per segment
  int vec_ok;      // outputs are aligned for vector stores
};

// kWarps: strips (warps) per CTA.  The warps of a CTA never meet, so the count
// only shapes the grid: the program's own choice (Prog::kWarps) for wide grids,
// one-warp CTAs for grids of a few strips, where every CTA is then full
// (C1: nine strips of a 2000-wide grid are three four-warp CTAs with a quarter
// of the warps idle; as nine one-warp CTAs the pass runs 26 % faster).
template <class Prog, int kWarps = Prog::kWarps>
struct Smem2D {
  // per warp: kStages slots, each holding kChunk rows of every input strip,
  // laid out [input][box][row][cell]: a TMA box is at most 256 elements wide,
  // so a 512-cell strip (16-cell lanes of 8- and 16-bit cells) arrives as two
  // boxes side by side, lanes 0-15 reading the first and 16-31 the second
  static constexpr int kStages = Prog::kStages;
  static constexpr int kChunk = Prog::kChunk;
  static constexpr int kStrip = Prog::kStrip;
  static constexpr int kBox0 = kStrip > 256 ? 256 : kStrip;
  static constexpr int kBoxes = kStrip / kBox0;
  static_assert(kStrip % kBox0 == 0 && kBox0 % Prog::kCells == 0,
                "strip is whole boxes, a lane stays inside one box");

  template <int M>
  __host__ __device__ static constexpr int row_bytes() {
    return int(sizeof(typename Prog::template T<M>)) * kBox0;
  }
  template <int M>
  __host__ __device__ static constexpr int box_bytes() {
    return row_bytes<M>() * kChunk;
  }
  template <int M>
  __host__ __device__ static constexpr int input_offset() {  // byte offset of input M in a slot
    if constexpr (M == 0) {
      return 0;
    } else {
      return input_offset<M - 1>() + box_bytes<M - 1>() * kBoxes;
    }
  }
  // byte offset of a lane's first cell in row 0 of input M
  template <int M>
  __device__ __forceinline__ static int lane_offset(int lane) {
    const int col = lane * Prog::kCells;
    if constexpr (kBoxes == 1) {
      return col * int(sizeof(typename Prog::template T<M>));
    } else {
      return (col / kBox0) * box_bytes<M>() +
             (col % kBox0) * int(sizeof(typename Prog::template T<M>));
    }
  }
  static constexpr int kSlotBytes = input_offset<Prog::kNumInputs>();
  static constexpr int kWarpBytes = kSlotBytes * kStages;
  static constexpr int kBarrierOffset = kWarpBytes * kWarps;
  static constexpr int kBytes =
      kBarrierOffset + int(sizeof(Mbarrier)) * kStages * kWarps;
};

template <class Prog>
struct Ctx2D {
  static constexpr bool kRotate = false;  // windows are shifted every step
  static constexpr int kPhase = 0;
  Rings<Prog> rings;
  const Params2D<Prog>& p;
  const unsigned char* slot_base;  // current slot of the TMA ring
  int lane;
  int x0;      // dimension-0 cell of the strip's first (lane 0) cell
  int seg_lo;  // output rows [seg_lo, seg_hi) belong to this warp
  int seg_hi;
  StorePlan store[Prog::kNumOutputs];
  bool any_partial;

  __device__ __forceinline__ explicit Ctx2D(const Params2D<Prog>& params)
      : p(params) {}
};

template <class Prog, int O = 0>
__device__ __forceinline__ void init_stores_2d(Ctx2D<Prog>& ctx, int t_begin) {
  if constexpr (O < Prog::kNumOutputs) {
    constexpr int N = Prog::kOutputNode[O];
    using T = typename Prog::template T<N>;
    constexpr int kC = Prog::kCells;
    const int col_in_strip = ctx.lane * kC;
    const bool lane_ok = col_in_strip >= Prog::kHaloLo0 &&
                         col_in_strip + kC <= Prog::kHaloLo0 + Prog::kValid0;
    init_store_plan<T, kC>(
        ctx.store[O], ctx.p.out[O], 0, ctx.p.out_pitch[O],
        ctx.x0 + col_in_strip, lane_ok, ctx.p.box_lo[O][0], ctx.p.box_hi[O][0],
        ctx.p.vec_ok != 0, max(ctx.seg_lo, ctx.p.box_lo[O][1]),
        min(ctx.seg_hi, ctx.p.box_hi[O][1]), t_begin - Prog::kNodes[N].lag);
    init_stores_2d<Prog, O + 1>(ctx, t_begin);
  }
}

template <class Prog, class Ctx, int N>
__device__ __forceinline__ void store_node_2d(Ctx& ctx, int t) {
  using T = typename Prog::template T<N>;
  constexpr int kOut = Prog::kNodes[N].out;
  T cells[Prog::kCells];
  newest_cells<N, 0, Prog>(ctx, cells);
  store_slice<T, Prog::kCells>(ctx.store[kOut], cells,
                               t - Prog::kNodes[N].lag, ctx.any_partial);
}

// One step: every node of the pass DAG produces one row.  `r` is the row of
// the current chunk that holds input row t.
template <class Prog, class Ctx, int N>
__device__ __forceinline__ void step_node_2d(Ctx& ctx, int t, int r) {
  using T = typename Prog::template T<N>;
  advance_ring<N, Prog>(ctx);
  if constexpr (Prog::kNodes[N].kind == 0) {
    constexpr int M = Prog::kNodes[N].src;
    using S = Smem2D<Prog>;
    const T* row = reinterpret_cast<const T*>(
        ctx.slot_base + S::template input_offset<M>() +
        r * S::template row_bytes<M>() + S::template lane_offset<M>(ctx.lane));
    T cells[Prog::kCells];
    load_shared_vec<T, Prog::kCells>(cells, row);
    units_from_cells<Prog, N>(newest_units<N, 0, Prog>(ctx), cells);
  } else {
    eval_units<Prog, Ctx, N>(ctx);
  }
  if constexpr (Prog::kNodes[N].out >= 0) store_node_2d<Prog, Ctx, N>(ctx, t);
}

// Producers first (kSkew == 0), or consumers first (kSkew == 1, pipelined
// plans): every node then reads only what earlier steps left in the windows,
// the nodes of a step are independent of each other and their dependent
// instruction chains overlap.
template <class Prog, class Ctx, int K = 0>
__device__ __forceinline__ void step_nodes_2d(Ctx& ctx, int t, int r) {
  if constexpr (K < Prog::kNumNodes) {
    constexpr int N = Prog::kSkew ? Prog::kNumNodes - 1 - K : K;
    step_node_2d<Prog, Ctx, N>(ctx, t, r);
    step_nodes_2d<Prog, Ctx, K + 1>(ctx, t, r);
  }
}

template <class Prog, int M = 0>
__device__ __forceinline__ void issue_chunk_2d(const Params2D<Prog>& p,
                                               unsigned char* slot, int x0,
                                               int row, Mbarrier* bar) {
  if constexpr (M < Prog::kNumInputs) {
    using S = Smem2D<Prog>;
#pragma unroll
    for (int b = 0; b < S::kBoxes; ++b) {
      tma_load_2d(slot + S::template input_offset<M>() +
                      b * S::template box_bytes<M>(),
                  &p.in_map[M], x0 + b * S::kBox0, row, bar);
    }
    issue_chunk_2d<Prog, M + 1>(p, slot, x0, row, bar);
  }
}

template <class Prog, int kWarps = Prog::kWarps>
__global__ void __launch_bounds__(kWarps * 32, Prog::kMinBlocks)
    soda_stream2d_kernel(const __grid_constant__ Params2D<Prog> p) {
  using S = Smem2D<Prog, kWarps>;
  constexpr int kStages = S::kStages;
  constexpr int kChunk = S::kChunk;
  unsigned char* smem = dyn_smem();
  const int warp = threadIdx.x >> 5;
  const int strip = blockIdx.x * kWarps + warp;
  if (strip >= p.num_strips) return;  // warps never meet at a CTA barrier

  Ctx2D<Prog> ctx(p);
  ctx.lane = lane_id();
  ctx.x0 = p.x_origin + strip * Prog::kValid0 - Prog::kHaloLo0;
  ctx.seg_lo = p.row_lo + blockIdx.y * p.seg_rows;
  ctx.seg_hi = min(ctx.seg_lo + p.seg_rows, p.row_hi);
  clear_rings<Prog>(ctx.rings);

  unsigned char* slots = smem + warp * S::kWarpBytes;
  Mbarrier* full =
      reinterpret_cast<Mbarrier*>(smem + S::kBarrierOffset) + warp * kStages;

  // input rows [t_begin, t_end] feed output rows [seg_lo, seg_hi)
  const int t_begin = ctx.seg_lo + Prog::kLoS;
  const int t_end = ctx.seg_hi - 1 + Prog::kMaxLag;
  const int num_chunks = (t_end - t_begin + kChunk) / kChunk;

  if (ctx.lane == 0) {
#pragma unroll
    for (int m = 0; m < Prog::kNumInputs; ++m) tma_prefetch_desc(&p.in_map[m]);
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    for (int s = 0; s < kStages && s < num_chunks; ++s) {
      mbar_arrive_expect_tx(&full[s], S::kSlotBytes);
      issue_chunk_2d<Prog>(p, slots + s * S::kSlotBytes, ctx.x0,
                           t_begin + s * kChunk, &full[s]);
    }
  }
  init_stores_2d<Prog>(ctx, t_begin);
  {
    bool partial = false;
#pragma unroll
    for (int o = 0; o < Prog::kNumOutputs; ++o)
      partial = partial || ctx.store[o].mode == 2;
    ctx.any_partial = warp_any(partial);
  }
  warp_sync();

  int slot = 0;
  unsigned parity = 0;
  for (int chunk = 0; chunk < num_chunks; ++chunk) {
    mbar_wait(&full[slot], parity);
    ctx.slot_base = slots + slot * S::kSlotBytes;
    const int t0 = t_begin + chunk * kChunk;
    if constexpr (Prog::kRowUnroll > 1) {
#pragma unroll
      for (int r = 0; r < kChunk; ++r) step_nodes_2d<Prog>(ctx, t0 + r, r);
    } else {
      // deep windows (contrast: 17 rows of 197 taps): one copy of the step body,
      // the window is shifted; unrolling would spill (measured: 12 KB of spill
      // loads per thread and 2x the time)
#pragma unroll 1
      for (int r = 0; r < kChunk; ++r) step_nodes_2d<Prog>(ctx, t0 + r, r);
    }
    warp_sync();  // every lane is done reading the slot
    if (ctx.lane == 0 && chunk + kStages < num_chunks) {
      mbar_arrive_expect_tx(&full[slot], S::kSlotBytes);
      issue_chunk_2d<Prog>(p, slots + slot * S::kSlotBytes, ctx.x0,
                           t_begin + (chunk + kStages) * kChunk, &full[slot]);
    }
    if (++slot == kStages) {
      slot = 0;
      parity ^= 1u;
    }
  }
}

// =============================================================================
// 3-D: a CTA streams a (strip x rows) tile along the last dimension.  Warp w
// owns the kCy tile rows [w * kCy, (w + 1) * kCy); a thread therefore holds a
// patch of kCy rows x kCells cells of every node in its register windows.
//   dimension-0 neighbours: warp shuffles (as in 2-D)
//   dimension-1 neighbours: registers inside the patch, shared-memory planes
//     outside it - the TMA ring itself for pass inputs, an export ring for
//     fused stages, read one step after they were written so that one CTA
//     barrier per step orders exports, their readers and the TMA slot reuse
//   streamed dimension: register windows
// The step loop is unrolled kUnroll times; ring depths that divide kUnroll get
// compile-time slots and window rotation by register renaming.
// =============================================================================

template <class Prog>
struct Params3D {
  TensorMap in_map[Prog::kNumInputs];
  void* out[Prog::kNumOutputs];
  long long out_pitch[Prog::kNumOutputs];        // elements between rows
  long long out_plane_pitch[Prog::kNumOutputs];  // elements between planes
  int box_lo[Prog::kNumOutputs][3];
  int box_hi[Prog::kNumOutputs][3];
  int x_origin;   // dim-0 cell of tile column 0's first valid cell
  int y_origin;   // dim-1 cell of tile row 0's first valid cell
  int plane_lo;   // first output plane produced by segment 0
  int plane_hi;
  int seg_planes;
  int vec_ok;
};

template <class Prog>
struct Smem3D {
  static constexpr int kStages = Prog::kStages;
  static constexpr int kPlaneCells = Prog::kStrip * Prog::kRows;

  template <int N>
  __host__ __device__ static constexpr int plane_bytes() {
    return int(sizeof(typename Prog::template T<N>)) * kPlaneCells;
  }
  // inputs are nodes 0 .. kNumInputs-1
  template <int M>
  __host__ __device__ static constexpr int input_offset() {
    if constexpr (M == 0) {
      return 0;
    } else {
      return input_offset<M - 1>() + plane_bytes<M - 1>();
    }
  }
  static constexpr int kSlotBytes = input_offset<Prog::kNumInputs>();
  static constexpr int kRingOffset = Prog::kGuardBytes;
  static constexpr int kExportOffset = kRingOffset + kSlotBytes * kStages;

  template <int N>
  __host__ __device__ static constexpr int export_offset() {  // relative to kExportOffset
    if constexpr (N == 0) {
      return 0;
    } else {
      return export_offset<N - 1>() +
             (Prog::kNodes[N - 1].kind == 1
                  ? Prog::kNodes[N - 1].smem_depth * plane_bytes<N - 1>()
                  : 0);
    }
  }
  static constexpr int kExportBytes = export_offset<Prog::kNumNodes>();
  static constexpr int kBarrierOffset =
      (kExportOffset + kExportBytes + Prog::kGuardBytes + 127) / 128 * 128;
  static constexpr int kBytes = kBarrierOffset + int(sizeof(Mbarrier)) * kStages;
};

// Where the cells of a lane's vector sit in a shared-memory plane.  TMA writes
// input planes in grid order; exported planes of packed programs keep the
// register order (lo, hi of unit 0, lo, hi of unit 1, ...) so that exports and
// aligned reloads are plain vector moves.
template <class Prog, int P>
__host__ __device__ constexpr int plane_index(int local) {
  if (Prog::kPack == 2 && Prog::kNodes[P].kind == 1) {
    return (local % kUnitsOf<Prog>) * 2 + local / kUnitsOf<Prog>;
  }
  return local;
}

template <class Prog>
struct Ctx3D {
  const Params3D<Prog>& p;
  unsigned char* smem;
  int lane;
  int row;       // first tile row of the thread's patch (warp * kCy)
  int cell_off;  // row * kStrip + lane * kCells
  int step;      // steps since the start of the segment (start of the round)
  int x0, y0;
  int seg_lo, seg_hi;
  // stores: one plan per output for patch row 0; the other rows are
  // out_pitch apart and only differ in whether they are stored at all
  StorePlan store[Prog::kNumOutputs];
  unsigned row_mask[Prog::kNumOutputs];
  bool any_partial;

  __device__ __forceinline__ explicit Ctx3D(const Params3D<Prog>& params)
      : p(params) {}
};

// What the functors see during step (round start + U) of the unrolled loop.
template <class Prog, int U>
struct Step3D {
  static constexpr bool kRotate = true;
  static constexpr int kPhase = U;
  Rings<Prog>& rings;
  Ctx3D<Prog>& c;

  // slot of a ring of kDepth slots that held step (current - kDistance)
  template <int kDepth, int kDistance>
  __device__ __forceinline__ int slot() const {
    if constexpr (Prog::kUnroll % kDepth == 0) {
      return ((U - kDistance) % kDepth + kDepth) % kDepth;
    } else {
      return (c.step + U + kDepth * (kDistance / kDepth + 1) - kDistance) %
             kDepth;
    }
  }

  // Plane of node P produced kDistance steps before P's current one.
  template <int P, int kDistance>
  __device__ __forceinline__ const typename Prog::template T<P>* plane() const {
    using S = Smem3D<Prog>;
    using T = typename Prog::template T<P>;
    if constexpr (Prog::kNodes[P].kind == 0) {
      return reinterpret_cast<const T*>(
          c.smem + S::kRingOffset +
          slot<S::kStages, kDistance>() * S::kSlotBytes +
          S::template input_offset<Prog::kNodes[P].src>());
    } else {
      constexpr int kDepth = Prog::kNodes[P].smem_depth;
      return reinterpret_cast<const T*>(
          c.smem + S::kExportOffset + S::template export_offset<P>() +
          slot<kDepth, kDistance>() * S::template plane_bytes<P>());
    }
  }

  // the vector of the lane kLaneOff lanes away, patch row kRow (any sign)
  template <int P, int kDistance, int kRow, int kLaneOff>
  __device__ __forceinline__ Vec<typename Prog::template T<P>, Prog::kCells>
  plane_vec() const {
    using T = typename Prog::template T<P>;
    return *reinterpret_cast<const Vec<T, Prog::kCells>*>(
        plane<P, kDistance>() + c.cell_off + kRow * Prog::kStrip +
        kLaneOff * Prog::kCells);
  }

  template <int P, int kDistance, int kCol, int kRow>
  __device__ __forceinline__ typename Prog::template T<P> plane_cell() const {
    constexpr int kLane = floor_div(kCol, Prog::kCells);
    constexpr int kLocal = kCol - kLane * Prog::kCells;
    return plane_vec<P, kDistance, kRow, kLane>()
        .v[plane_index<Prog, P>(kLocal)];
  }

  // packed: cells (kCol, kCol + kCells / 2)
  template <int P, int kDistance, int kCol, int kRow>
  __device__ __forceinline__ auto plane_unit() const {
    constexpr int kHalf = kUnitsOf<Prog>;
    return pair_pack(plane_cell<P, kDistance, kCol, kRow>(),
                     plane_cell<P, kDistance, kCol + kHalf, kRow>());
  }
};

template <class Prog, int O = 0>
__device__ __forceinline__ void init_stores_3d(Ctx3D<Prog>& ctx, int t_begin) {
  if constexpr (O < Prog::kNumOutputs) {
    constexpr int N = Prog::kOutputNode[O];
    using T = typename Prog::template T<N>;
    constexpr int kC = Prog::kCells;
    const int col_in_strip = ctx.lane * kC;
    const int y = ctx.y0 + ctx.row;  // patch row 0
    unsigned rows = 0;
#pragma unroll
    for (int j = 0; j < Prog::kCy; ++j) {
      const int tile_row = ctx.row + j;
      if (tile_row >= Prog::kHaloLo1 &&
          tile_row < Prog::kHaloLo1 + Prog::kValid1 &&
          y + j >= ctx.p.box_lo[O][1] && y + j < ctx.p.box_hi[O][1])
        rows |= 1u << j;
    }
    const bool lane_ok = col_in_strip >= Prog::kHaloLo0 &&
                         col_in_strip + kC <= Prog::kHaloLo0 + Prog::kValid0 &&
                         rows != 0;
    init_store_plan<T, kC>(
        ctx.store[O], ctx.p.out[O],
        static_cast<long long>(y) * ctx.p.out_pitch[O], ctx.p.out_plane_pitch[O],
        ctx.x0 + col_in_strip, lane_ok, ctx.p.box_lo[O][0], ctx.p.box_hi[O][0],
        ctx.p.vec_ok != 0, max(ctx.seg_lo, ctx.p.box_lo[O][2]),
        min(ctx.seg_hi, ctx.p.box_hi[O][2]), t_begin - Prog::kNodes[N].lag);
    ctx.row_mask[O] = ctx.store[O].mode == 0 ? 0u : rows;
    init_stores_3d<Prog, O + 1>(ctx, t_begin);
  }
}

template <class Prog, class Step, int N, int J = 0>
__device__ __forceinline__ void store_rows_3d(Step& st, unsigned char* ptr) {
  if constexpr (J < Prog::kCy) {
    using T = typename Prog::template T<N>;
    constexpr int kOut = Prog::kNodes[N].out;
    constexpr int kC = Prog::kCells;
    const Ctx3D<Prog>& ctx = st.c;
    const StorePlan& plan = ctx.store[kOut];
    if ((ctx.row_mask[kOut] >> J) & 1u) {
      T cells[kC];
      newest_cells<N, J, Prog>(st, cells);
      T* dst = reinterpret_cast<T*>(ptr) + J * ctx.p.out_pitch[kOut];
      if (!ctx.any_partial) {
        store_global_vec<T, kC>(dst, cells);
      } else {
#pragma unroll
        for (int i = 0; i < kC; ++i) {
          if ((plan.mask >> i) & 1u) dst[i] = cells[i];
        }
      }
    }
    store_rows_3d<Prog, Step, N, J + 1>(st, ptr);
  }
}

template <class Prog, class Step, int N>
__device__ __forceinline__ void store_node_3d(Step& st, int t) {
  constexpr int kOut = Prog::kNodes[N].out;
  StorePlan& plan = st.c.store[kOut];
  unsigned char* ptr = plan.ptr;
  plan.ptr += plan.step;
  const int slice = t - Prog::kNodes[N].lag;
  if (slice < plan.slice_lo || slice >= plan.slice_hi) return;  // uniform
  // without partial lanes in the warp, mode is 0 (row_mask == 0) or 1
  store_rows_3d<Prog, Step, N>(st, ptr);
}

// rows of the patch that other warps read: the kReach rows next to each edge
template <class Prog, class Step, int N, int J = 0>
__device__ __forceinline__ void export_rows_3d(Step& st) {
  if constexpr (J < Prog::kCy) {
    constexpr int kReach = Prog::kNodes[N].smem_reach;
    if constexpr (J < kReach || J >= Prog::kCy - kReach) {
      using T = typename Prog::template T<N>;
      constexpr int kC = Prog::kCells;
      T* dst = const_cast<T*>(st.template plane<N, 0>()) + st.c.cell_off +
               J * Prog::kStrip;
      Vec<T, kC> tmp;
      if constexpr (Prog::kPack == 2) {
        const auto& units = newest_units<N, J, Prog>(st);
#pragma unroll
        for (int u = 0; u < kUnitsOf<Prog>; ++u) {
          tmp.v[2 * u] = pair_lo(units[u]);
          tmp.v[2 * u + 1] = pair_hi(units[u]);
        }
      } else {
        T cells[kC];
        newest_cells<N, J, Prog>(st, cells);
#pragma unroll
        for (int i = 0; i < kC; ++i) tmp.v[i] = cells[i];
      }
      *reinterpret_cast<Vec<T, kC>*>(dst) = tmp;
    }
    export_rows_3d<Prog, Step, N, J + 1>(st);
  }
}

template <class Prog, class Step, int N, int J = 0>
__device__ __forceinline__ void load_input_rows_3d(Step& st) {
  if constexpr (J < Prog::kCy) {
    using T = typename Prog::template T<N>;
    T cells[Prog::kCells];
    load_shared_vec<T, Prog::kCells>(
        cells, st.template plane<N, 0>() + st.c.cell_off + J * Prog::kStrip);
    units_from_cells<Prog, N>(newest_units<N, J, Prog>(st), cells);
    load_input_rows_3d<Prog, Step, N, J + 1>(st);
  }
}

template <class Prog, class Step, int N = 0>
__device__ __forceinline__ void step_nodes_3d(Step& st, int t) {
  if constexpr (N < Prog::kNumNodes) {
    advance_ring<N, Prog>(st);
    if constexpr (Prog::kNodes[N].kind == 0) {
      load_input_rows_3d<Prog, Step, N>(st);
    } else {
      eval_units<Prog, Step, N>(st);
      // export for the dimension-1 neighbours (read from the next step on)
      if constexpr (Prog::kNodes[N].smem_depth > 0)
        export_rows_3d<Prog, Step, N>(st);
    }
    if constexpr (Prog::kNodes[N].out >= 0)
      store_node_3d<Prog, Step, N>(st, t);
    step_nodes_3d<Prog, Step, N + 1>(st, t);
  }
}

template <class Prog, int M = 0>
__device__ __forceinline__ void issue_plane_3d(const Params3D<Prog>& p,
                                               unsigned char* slot, int x0,
                                               int y0, int plane,
                                               Mbarrier* bar) {
  if constexpr (M < Prog::kNumInputs) {
    tma_load_3d(slot + Smem3D<Prog>::template input_offset<M>(), &p.in_map[M],
                x0, y0, plane, bar);
    issue_plane_3d<Prog, M + 1>(p, slot, x0, y0, plane, bar);
  }
}

// steps [round, round + kUnroll) of the segment.  There is no per-step guard:
// the kernel rounds the number of steps up to a multiple of kUnroll (planes
// past the end of the grid are zero-filled by TMA and nothing is stored for
// them), so that the register windows rotate by renaming across the round.
template <class Prog, int U = 0>
__device__ __forceinline__ void round_3d(Rings<Prog>& rings, Ctx3D<Prog>& ctx,
                                         Mbarrier* full, int t_begin,
                                         int num_steps) {
  if constexpr (U < Prog::kUnroll) {
    using S = Smem3D<Prog>;
    constexpr int kStages = S::kStages;
    constexpr int kInDepth = Prog::kInDepth;  // input planes still readable
    const int step = ctx.step + U;
    Step3D<Prog, U> st{rings, ctx};
    const int slot = st.template slot<kStages, 0>();
    mbar_wait(&full[slot], (static_cast<unsigned>(step) / kStages) & 1u);
    step_nodes_3d<Prog>(st, t_begin + step);
    cta_sync();  // exports visible; plane (step - kInDepth + 1) is dead
    if (threadIdx.x < 32) {  // warp-uniform
      const int dead = step - (kInDepth - 1);
      if (ctx.lane == 0 && dead >= 0 && dead + kStages < num_steps) {
        const int s = st.template slot<kStages, kInDepth - 1>();
        mbar_arrive_expect_tx(&full[s], S::kSlotBytes);
        issue_plane_3d<Prog>(ctx.p, ctx.smem + S::kRingOffset + s * S::kSlotBytes,
                             ctx.x0, ctx.y0, t_begin + dead + kStages, &full[s]);
      }
    }
    round_3d<Prog, U + 1>(rings, ctx, full, t_begin, num_steps);
  }
}

template <class Prog>
__global__ void __launch_bounds__(Prog::kWarps * 32, Prog::kMinBlocks)
    soda_stream3d_kernel(const __grid_constant__ Params3D<Prog> p) {
  using S = Smem3D<Prog>;
  constexpr int kStages = S::kStages;
  static_assert(kStages > Prog::kInDepth, "TMA ring needs look-ahead slots");
  static_assert(Prog::kRows == Prog::kWarps * Prog::kCy, "tile rows");

  Rings<Prog> rings;
  Ctx3D<Prog> ctx(p);
  ctx.smem = dyn_smem();
  ctx.lane = lane_id();
  ctx.row = (threadIdx.x >> 5) * Prog::kCy;
  ctx.cell_off = ctx.row * Prog::kStrip + ctx.lane * Prog::kCells;
  ctx.x0 = p.x_origin + blockIdx.x * Prog::kValid0 - Prog::kHaloLo0;
  ctx.y0 = p.y_origin + blockIdx.y * Prog::kValid1 - Prog::kHaloLo1;
  ctx.seg_lo = p.plane_lo + blockIdx.z * p.seg_planes;
  ctx.seg_hi = min(ctx.seg_lo + p.seg_planes, p.plane_hi);
  clear_rings<Prog>(rings);

  Mbarrier* full = reinterpret_cast<Mbarrier*>(ctx.smem + S::kBarrierOffset);
  unsigned char* ring = ctx.smem + S::kRingOffset;
  const int t_begin = ctx.seg_lo + Prog::kLoS;
  const int num_steps =
      (ctx.seg_hi - 1 + Prog::kMaxLag - t_begin + Prog::kUnroll) /
      Prog::kUnroll * Prog::kUnroll;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int m = 0; m < Prog::kNumInputs; ++m) tma_prefetch_desc(&p.in_map[m]);
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    for (int s = 0; s < kStages && s < num_steps; ++s) {
      mbar_arrive_expect_tx(&full[s], S::kSlotBytes);
      issue_plane_3d<Prog>(p, ring + s * S::kSlotBytes, ctx.x0, ctx.y0,
                           t_begin + s, &full[s]);
    }
  }
  init_stores_3d<Prog>(ctx, t_begin);
  {
    bool partial = false;
#pragma unroll
    for (int o = 0; o < Prog::kNumOutputs; ++o)
      partial = partial || ctx.store[o].mode == 2;
    ctx.any_partial = warp_any(partial);
  }
  cta_sync();

  for (ctx.step = 0; ctx.step < num_steps; ctx.step += Prog::kUnroll)
    round_3d<Prog>(rings, ctx, full, t_begin, num_steps);
}

}  // namespace soda
