// GPU tiler / un-tiler for the reference's stream data layout (C ABI:
// include/soda_layout.h).  Restates the two loop nests of the generated host
// wrapper soda::app::<app>() as gather kernels:
//   pack    /root/reference/src/soda/codegen/frt/host.py:181-249
//   unpack  /root/reference/src/soda/codegen/frt/host.py:340-427
// Both are pure data movement: the roofline is HBM, the algorithmic traffic
// one read + one write per stream element (pack) or per valid cell (unpack).
// A thread walks kRun consecutive elements so that the div/mod chain that
// turns a stream offset into (tile, coordinates) is paid once per run and not
// once per element (it would otherwise cost more issue slots than the bytes
// cost HBM time).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <string>

#include "soda_layout.h"

namespace {

// consecutive elements per thread: 32 bytes' worth, at least 8
template <typename T>
constexpr int kRunOf = sizeof(T) >= 4 ? 8 : 32 / static_cast<int>(sizeof(T)) > 16 ? 16 : 32 / static_cast<int>(sizeof(T));
constexpr int kThreads = 256;

thread_local std::string g_error;
std::atomic<long long> g_launches{0};

int fail(int status, const std::string& message) {
  g_error = message;
  return status;
}

struct Geometry {
  int dim, banks;
  int bank_shift;  // log2(banks) when banks is a power of two, else -1
  int extent[3], tile_size[3], stencil_dim[3], window_offset[3], window_dim[3];
  int tile_count[3], tile_stride[3];
  long long stride[3];
  long long aligned, tiles_total, stencil_offset, produce_offset;
  long long bank_elems;
  void* bank[SODA_LAYOUT_MAX_BANKS];
};

struct Position {  // one stream offset, decoded
  int idx[2];      // tile index in dims 0, 1
  int c[3];        // coordinates in the tile
  long long off;   // offset inside the tile's aligned block
};

__device__ __forceinline__ int actual_tile_size(const Geometry& g, int d, int index) {
  return index == g.tile_count[d] - 1 ? g.extent[d] - g.tile_stride[d] * index
                                      : g.tile_size[d];
}

// stream offset -> tile and coordinates; false: beyond the last tile
__device__ __forceinline__ bool decode(const Geometry& g, long long t, Position* p) {
  const long long tile = t / g.aligned;
  if (tile >= g.tiles_total) return false;
  p->off = t - tile * g.aligned;
  p->idx[0] = static_cast<int>(tile % g.tile_count[0]);
  p->idx[1] = g.dim == 3 ? static_cast<int>(tile / g.tile_count[0]) : 0;
  long long rest = p->off;
  p->c[0] = static_cast<int>(rest % g.tile_size[0]);
  rest /= g.tile_size[0];
  if (g.dim == 3) {
    p->c[1] = static_cast<int>(rest % g.tile_size[1]);
    rest /= g.tile_size[1];
  } else {
    p->c[1] = 0;
  }
  // padding at the end of the aligned block decodes to a last-dimension
  // coordinate >= extent: void
  p->c[2] = rest > 0x7fffffff ? 0x7fffffff : static_cast<int>(rest);
  return true;
}

__device__ __forceinline__ bool inside(const Geometry& g, const Position& p) {
  if (p.c[2] >= g.extent[g.dim - 1]) return false;
  if (p.c[0] >= actual_tile_size(g, 0, p.idx[0])) return false;
  if (g.dim == 3 && p.c[1] >= actual_tile_size(g, 1, p.idx[1])) return false;
  return true;
}

__device__ __forceinline__ long long original_offset(const Geometry& g,
                                                     const Position& p) {
  long long o = (static_cast<long long>(p.idx[0]) * g.tile_stride[0] + p.c[0]) *
                g.stride[0];
  if (g.dim == 3) {
    o += (static_cast<long long>(p.idx[1]) * g.tile_stride[1] + p.c[1]) *
         g.stride[1];
    o += static_cast<long long>(p.c[2]) * g.stride[2];
  } else {
    o += static_cast<long long>(p.c[2]) * g.stride[1];
  }
  return o;
}

// (bank, position in the bank) of a stream offset, advanced without dividing
struct BankCursor {
  int bank;
  long long pos;
  __device__ __forceinline__ BankCursor(const Geometry& g, long long t) {
    if (g.bank_shift >= 0) {
      bank = static_cast<int>(t & (g.banks - 1));
      pos = t >> g.bank_shift;
    } else {
      bank = static_cast<int>(t % g.banks);
      pos = t / g.banks;
    }
  }
  __device__ __forceinline__ void next(const Geometry& g) {
    if (++bank == g.banks) {
      bank = 0;
      ++pos;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
    pack_kernel(const Geometry g, const T* __restrict__ dense) {
  constexpr int kRun = kRunOf<T>;
  const long long total = g.bank_elems * g.banks;
  long long t = (static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x) * kRun;
  if (t >= total) return;
  Position p;
  bool in_tiles = decode(g, t, &p);
  BankCursor out(g, t);
  // fast path: the whole run lies in one row of one tile (the usual case)
  if (in_tiles && t + kRun <= total && p.off + kRun <= g.aligned &&
      p.c[0] + kRun <= g.tile_size[0]) {
    const bool row_ok = p.c[2] < g.extent[g.dim - 1] &&
                        (g.dim == 2 || p.c[1] < actual_tile_size(g, 1, p.idx[1]));
    const int width = row_ok ? actual_tile_size(g, 0, p.idx[0]) - p.c[0] : 0;
    const long long src = original_offset(g, p) - g.produce_offset;
    T v[kRun];
#pragma unroll
    for (int k = 0; k < kRun; ++k) {
      const long long at = src + k;
      v[k] = k < width ? dense[at > 0 ? at : 0] : T(0);
    }
    if (g.banks == 1) {
      T* dst = static_cast<T*>(g.bank[0]) + t;  // t is a multiple of kRun
      constexpr int kVec = 16 / sizeof(T) < kRun ? 16 / sizeof(T) : kRun;
      struct alignas(sizeof(T) * kVec) Pack { T e[kVec]; };
#pragma unroll
      for (int k = 0; k < kRun; k += kVec) {
        Pack pack;
#pragma unroll
        for (int i = 0; i < kVec; ++i) pack.e[i] = v[k + i];
        *reinterpret_cast<Pack*>(dst + k) = pack;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kRun; ++k) {
        static_cast<T*>(g.bank[out.bank])[out.pos] = v[k];
        out.next(g);
      }
    }
    return;
  }
  // slow path: rows, tiles or the stream end inside the run
  for (int k = 0; k < kRun && t < total; ++k, ++t) {
    T value = T(0);
    if (in_tiles && inside(g, p)) {
      long long src = original_offset(g, p) - g.produce_offset;
      value = dense[src > 0 ? src : 0];
    }
    static_cast<T*>(g.bank[out.bank])[out.pos] = value;
    out.next(g);
    if (in_tiles) {
      ++p.off;
      if (++p.c[0] == g.tile_size[0] || p.off == g.aligned)
        in_tiles = decode(g, t + 1, &p);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    unpack_kernel(const Geometry g, T* __restrict__ dense, int x_lo, int x_hi,
                  int y_lo, int y_hi, int z_lo, int z_hi) {
  constexpr int kRun = kRunOf<T>;
  // a warp walks 32 * kRun consecutive cells of one row of the valid box, lane
  // l taking cells l, l + 32, ...: loads and stores of a warp are contiguous
  constexpr int kSpan = 32 * kRun;
  const int spans_per_row = (x_hi - x_lo + kSpan - 1) / kSpan;
  const long long warp =
      (static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const long long rows = static_cast<long long>(y_hi - y_lo) * (z_hi - z_lo);
  if (warp >= rows * spans_per_row) return;
  const long long row = warp / spans_per_row;
  int x = x_lo + static_cast<int>(warp - row * spans_per_row) * kSpan +
          (threadIdx.x & 31);
  if (x >= x_hi) return;
  const int y = y_lo + static_cast<int>(row % (y_hi - y_lo));
  const int z = z_lo + static_cast<int>(row / (y_hi - y_lo));

  const int lo0 = g.window_offset[0] > 0 ? g.window_offset[0] : 0;
  const int cut0 = g.window_dim[0] - 1 - g.window_offset[0] > 0
                       ? g.window_dim[0] - 1 - g.window_offset[0] : 0;
  // tile that holds the valid copy of column x: the one whose valid range
  // [k S + lo, (k + 1) S + lo) contains it (the last tile takes the rest)
  const int last = g.tile_count[0] - 1;
  int idx0 = (x - lo0) / g.tile_stride[0];
  if (idx0 > last) idx0 = last;
  int c0 = x - idx0 * g.tile_stride[0];
  long long tile, in_tile, dst;
  if (g.dim == 3) {
    const int lo1 = g.window_offset[1] > 0 ? g.window_offset[1] : 0;
    int idx1 = (y - lo1) / g.tile_stride[1];
    if (idx1 > g.tile_count[1] - 1) idx1 = g.tile_count[1] - 1;
    const int c1 = y - idx1 * g.tile_stride[1];
    const int cut1 = g.window_dim[1] - 1 - g.window_offset[1] > 0
                         ? g.window_dim[1] - 1 - g.window_offset[1] : 0;
    // windows that start right of the origin leave rows no tile stores
    if (c1 >= actual_tile_size(g, 1, idx1) - cut1) return;
    tile = static_cast<long long>(idx1) * g.tile_count[0];
    in_tile = (static_cast<long long>(z) * g.tile_size[1] + c1) * g.tile_size[0];
    dst = y * g.stride[1] + z * g.stride[2];
  } else {
    tile = 0;
    in_tile = static_cast<long long>(y) * g.tile_size[0];
    dst = y * g.stride[1];
  }
  T* out = dense + dst;
  // stream offset of column c0 of the current tile is base + c0
  long long base = (tile + idx0) * g.aligned + in_tile + g.stencil_offset;
  int limit = actual_tile_size(g, 0, idx0) - cut0;  // first column not stored
  // fast path: all kRun cells of this lane lie in the valid range of one tile
  {
    const int c_last = c0 + 32 * (kRun - 1);
    const bool same_tile = idx0 == last || c_last < g.tile_stride[0] + lo0;
    if (same_tile && c_last < limit && x + 32 * (kRun - 1) < x_hi) {
      const long long t0 = base + c0;
      if (g.banks == 1) {
        const T* src = static_cast<const T*>(g.bank[0]) + t0;
        T v[kRun];
#pragma unroll
        for (int k = 0; k < kRun; ++k) v[k] = src[32 * k];
#pragma unroll
        for (int k = 0; k < kRun; ++k) out[x + 32 * k] = v[k];
      } else {
        T v[kRun];
#pragma unroll
        for (int k = 0; k < kRun; ++k) {
          const BankCursor in(g, t0 + 32 * k);
          v[k] = static_cast<const T*>(g.bank[in.bank])[in.pos];
        }
#pragma unroll
        for (int k = 0; k < kRun; ++k) out[x + 32 * k] = v[k];
      }
      return;
    }
  }
#pragma unroll 1
  for (int k = 0; k < kRun; ++k, x += 32, c0 += 32) {
    if (x >= x_hi) return;
    while (c0 >= g.tile_stride[0] + lo0 && idx0 < last) {
      ++idx0;  // the valid range of the next tile starts here
      c0 -= g.tile_stride[0];
      base += g.aligned;
      limit = actual_tile_size(g, 0, idx0) - cut0;
    }
    // a column between two tiles' valid ranges is not stored by the reference
    if (c0 >= limit) continue;
    const BankCursor in(g, base + c0);
    out[x] = static_cast<const T*>(g.bank[in.bank])[in.pos];
  }
}

int make_geometry(const soda_stream_layout* l, Geometry* g) {
  if (l == nullptr || l->struct_size != static_cast<int32_t>(sizeof(*l)))
    return fail(SODA_LAYOUT_BAD_ARGUMENT, "layout is NULL or has a wrong struct_size");
  if (l->dim != 2 && l->dim != 3)
    return fail(SODA_LAYOUT_UNSUPPORTED, "only 2-D and 3-D layouts");
  if (l->banks < 1 || l->banks > SODA_LAYOUT_MAX_BANKS)
    return fail(SODA_LAYOUT_BAD_ARGUMENT, "bad bank count");
  if (l->elem_bytes != 1 && l->elem_bytes != 2 && l->elem_bytes != 4 &&
      l->elem_bytes != 8)
    return fail(SODA_LAYOUT_UNSUPPORTED, "element size must be 1, 2, 4 or 8 bytes");
  if (l->stride[0] != 1)
    return fail(SODA_LAYOUT_UNSUPPORTED, "stride[0] must be 1 (dimension 0 dense)");
  if (l->elem_count_aligned_per_tile <= 0 || l->elem_count_per_cycle <= 0)
    return fail(SODA_LAYOUT_BAD_ARGUMENT, "bad alignment constants");
  memset(g, 0, sizeof(*g));
  g->dim = l->dim;
  g->banks = l->banks;
  g->bank_shift = -1;
  for (int k = 0; k < 6; ++k)
    if ((1 << k) == l->banks) g->bank_shift = k;
  g->tiles_total = 1;
  for (int d = 0; d < l->dim; ++d) {
    if (l->extent[d] <= 0)
      return fail(SODA_LAYOUT_BAD_ARGUMENT, "extent must be positive");
    g->extent[d] = l->extent[d];
    g->stride[d] = l->stride[d];
    g->stencil_dim[d] = l->stencil_dim[d];
    g->window_offset[d] = l->window_offset[d];
    g->window_dim[d] = l->window_dim[d];
  }
  for (int d = 0; d < l->dim - 1; ++d) {
    g->tile_size[d] = l->tile_size[d];
    g->tile_stride[d] = l->tile_size[d] - l->stencil_dim[d] + 1;
    if (g->tile_stride[d] <= 0)
      return fail(SODA_LAYOUT_BAD_ARGUMENT, "tile smaller than the stencil window");
    // host.py:124-128
    g->tile_count[d] = (l->extent[d] - l->stencil_dim[d] + 1 - 1) / g->tile_stride[d] + 1;
    if (l->extent[d] < l->stencil_dim[d] || g->tile_count[d] < 1)
      return fail(SODA_LAYOUT_BAD_ARGUMENT, "extent smaller than the stencil window");
    g->tiles_total *= g->tile_count[d];
  }
  g->aligned = l->elem_count_aligned_per_tile;
  long long per_tile = l->extent[l->dim - 1];
  for (int d = 0; d < l->dim - 1; ++d) per_tile *= l->tile_size[d];
  if (g->aligned < per_tile)
    return fail(SODA_LAYOUT_BAD_ARGUMENT,
                "elem_count_aligned_per_tile is smaller than a tile (inputs and "
                "outputs with different elements per cycle: the reference's "
                "buffers would overlap, host.py:138-145)");
  g->stencil_offset = l->stencil_offset;
  g->produce_offset = l->produce_offset;
  // host.py:147-162
  const long long epc = l->elem_count_per_cycle;
  const long long tail = ((l->stencil_distance - 1) / epc + 1) * epc;
  g->bank_elems = (g->tiles_total * g->aligned +
                   (l->stencil_distance > 0 ? tail : 0)) / l->banks;
  return SODA_LAYOUT_OK;
}

#define LAYOUT_CUDA_CHECK(expr)                                              \
  do {                                                                       \
    cudaError_t err_ = (expr);                                               \
    if (err_ != cudaSuccess)                                                 \
      return fail(SODA_LAYOUT_CUDA_ERROR,                                    \
                  std::string(#expr) + ": " + cudaGetErrorString(err_));     \
  } while (0)

template <typename T>
int launch_pack(const Geometry& g, const void* dense, cudaStream_t stream) {
  constexpr int kRun = kRunOf<T>;
  const long long total = g.bank_elems * g.banks;
  const long long threads = (total + kRun - 1) / kRun;
  const long long blocks = (threads + kThreads - 1) / kThreads;
  if (blocks == 0) return SODA_LAYOUT_OK;
  if (blocks > 0x7fffffffLL) return fail(SODA_LAYOUT_UNSUPPORTED, "grid too large");
  pack_kernel<T><<<static_cast<unsigned>(blocks), kThreads, 0, stream>>>(
      g, static_cast<const T*>(dense));
  g_launches.fetch_add(1);
  LAYOUT_CUDA_CHECK(cudaGetLastError());
  return SODA_LAYOUT_OK;
}

template <typename T>
int launch_unpack(const Geometry& g, void* dense, cudaStream_t stream) {
  constexpr int kRun = kRunOf<T>;
  // valid box of the whole grid (host.py:357-376 over all tiles)
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < g.dim; ++d) {
    const int cut = g.window_dim[d] - 1 - g.window_offset[d];
    lo[d] = g.window_offset[d] > 0 ? g.window_offset[d] : 0;
    hi[d] = g.extent[d] - (cut > 0 ? cut : 0);
    if (hi[d] <= lo[d]) return SODA_LAYOUT_OK;
  }
  int y_lo = lo[1], y_hi = hi[1], z_lo = 0, z_hi = 1;
  if (g.dim == 3) {
    z_lo = lo[2];
    z_hi = hi[2];
  }
  const long long spans = (hi[0] - lo[0] + 32 * kRun - 1) / (32 * kRun);
  const long long warps = spans * (y_hi - y_lo) * (z_hi - z_lo);
  const long long blocks = (warps * 32 + kThreads - 1) / kThreads;
  if (blocks > 0x7fffffffLL) return fail(SODA_LAYOUT_UNSUPPORTED, "grid too large");
  unpack_kernel<T><<<static_cast<unsigned>(blocks), kThreads, 0, stream>>>(
      g, static_cast<T*>(dense), lo[0], hi[0], y_lo, y_hi, z_lo, z_hi);
  g_launches.fetch_add(1);
  LAYOUT_CUDA_CHECK(cudaGetLastError());
  return SODA_LAYOUT_OK;
}

}  // namespace

extern "C" {

int soda_layout_bank_elems(const soda_stream_layout* layout, int64_t* elems) {
  Geometry g;
  int status = make_geometry(layout, &g);
  if (status != SODA_LAYOUT_OK) return status;
  if (elems == nullptr) return fail(SODA_LAYOUT_BAD_ARGUMENT, "elems is NULL");
  *elems = g.bank_elems;
  return SODA_LAYOUT_OK;
}

int soda_layout_pack_device(const soda_stream_layout* layout, const void* dense,
                            void* const* bank_buffers, void* stream) {
  Geometry g;
  int status = make_geometry(layout, &g);
  if (status != SODA_LAYOUT_OK) return status;
  if (dense == nullptr || bank_buffers == nullptr)
    return fail(SODA_LAYOUT_BAD_ARGUMENT, "NULL buffer");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(SODA_LAYOUT_CUDA_ERROR, "no CUDA device: there is no CPU path");
  for (int b = 0; b < g.banks; ++b) {
    if (bank_buffers[b] == nullptr)
      return fail(SODA_LAYOUT_BAD_ARGUMENT, "NULL bank buffer");
    g.bank[b] = bank_buffers[b];
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (layout->elem_bytes) {
    case 1: return launch_pack<uint8_t>(g, dense, s);
    case 2: return launch_pack<uint16_t>(g, dense, s);
    case 4: return launch_pack<uint32_t>(g, dense, s);
    default: return launch_pack<uint64_t>(g, dense, s);
  }
}

int soda_layout_unpack_device(const soda_stream_layout* layout,
                              const void* const* bank_buffers, void* dense,
                              void* stream) {
  Geometry g;
  int status = make_geometry(layout, &g);
  if (status != SODA_LAYOUT_OK) return status;
  if (dense == nullptr || bank_buffers == nullptr)
    return fail(SODA_LAYOUT_BAD_ARGUMENT, "NULL buffer");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(SODA_LAYOUT_CUDA_ERROR, "no CUDA device: there is no CPU path");
  for (int b = 0; b < g.banks; ++b) {
    if (bank_buffers[b] == nullptr)
      return fail(SODA_LAYOUT_BAD_ARGUMENT, "NULL bank buffer");
    g.bank[b] = const_cast<void*>(bank_buffers[b]);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (layout->elem_bytes) {
    case 1: return launch_unpack<uint8_t>(g, dense, s);
    case 2: return launch_unpack<uint16_t>(g, dense, s);
    case 4: return launch_unpack<uint32_t>(g, dense, s);
    default: return launch_unpack<uint64_t>(g, dense, s);
  }
}

int64_t soda_layout_launch_count(void) { return g_launches.load(); }

const char* soda_layout_last_error(void) { return g_error.c_str(); }

}  // extern "C"
