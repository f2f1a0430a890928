/* C ABI of the SODA stream-data-layout codec (GPU pack / unpack kernels).
 *
 * The reference's generated host wrapper soda::app::<app>() does not hand user
 * arrays to the kernel directly: it *tiles* every input into per-bank,
 * burst-aligned stream buffers (overlapping tiles of tile_size cells in every
 * dimension but the last, cyclic partition over the DRAM banks, void padding
 * up to the burst width and kStencilDistance void elements at the end) and
 * *un-tiles* the kernel's output buffers back into the valid interior of the
 * user's output array:
 *   tiler     /root/reference/src/soda/codegen/frt/host.py:181-249
 *   un-tiler  /root/reference/src/soda/codegen/frt/host.py:340-427
 *   buffers   /root/reference/src/soda/codegen/frt/host.py:112-179
 *   pictures  /root/reference/docs/data-layout.md
 * These entry points do the same two transformations on the GPU, on device
 * pointers, so that a buffer in the reference's stream format (a file written
 * for / read from an FPGA run) can be exchanged with a dense device array
 * without a host round trip.  They replace the two loop nests named above and
 * nothing else; see INTEGRATION.md for the binding.
 */
#ifndef SODA_LAYOUT_H_
#define SODA_LAYOUT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SODA_LAYOUT_API __declspec(dllexport)
#else
#define SODA_LAYOUT_API __attribute__((visibility("default")))
#endif

#define SODA_LAYOUT_MAX_DIM 3
#define SODA_LAYOUT_MAX_BANKS 32

enum soda_layout_status {
  SODA_LAYOUT_OK = 0,
  SODA_LAYOUT_BAD_ARGUMENT = 1,
  SODA_LAYOUT_CUDA_ERROR = 2,
  SODA_LAYOUT_UNSUPPORTED = 3,
};

/* Everything the reference's loops read, for ONE tensor on ONE grid.
 * Names follow the generated C++ (frt/host.py). */
typedef struct soda_stream_layout {
  int32_t struct_size;
  int32_t dim;                               /* 2 or 3 */
  int32_t elem_bytes;                        /* 1, 2, 4 or 8 */
  int32_t banks;                             /* bank_count_<name> */
  int32_t extent[SODA_LAYOUT_MAX_DIM];       /* var_<name>_extent */
  int64_t stride[SODA_LAYOUT_MAX_DIM];       /* var_<name>_stride, elements; stride[0] == 1 */
  int32_t tile_size[SODA_LAYOUT_MAX_DIM];    /* tile_size_<d>, d < dim-1 */
  int32_t stencil_dim[SODA_LAYOUT_MAX_DIM];  /* kStencilDim<d> */
  /* un-tiler loop bounds: offset and size of the window first input -> first
   * output (host.py:352-376) */
  int32_t window_offset[SODA_LAYOUT_MAX_DIM];
  int32_t window_dim[SODA_LAYOUT_MAX_DIM];
  int64_t stencil_distance;                  /* kStencilDistance */
  int64_t stencil_offset;                    /* output: distance - serialize(offset); input: 0 */
  int64_t produce_offset;                    /* input: tensors[x].produce_offset; output: 0 */
  int64_t elem_count_aligned_per_tile;       /* ..._i for inputs, ..._o for outputs */
  int64_t elem_count_per_cycle;              /* burst_width / width * banks of this tensor */
} soda_stream_layout;

/* Elements every bank buffer of the tensor holds (buf_size_<name> /
 * sizeof(T), host.py:147-162). */
SODA_LAYOUT_API int soda_layout_bank_elems(const soda_stream_layout* layout,
                                           int64_t* elems);

/* Tiler: dense device array -> `banks` device buffers of
 * soda_layout_bank_elems() elements each.  Every element of every buffer is
 * written; positions the reference leaves uninitialised ("void") become 0.
 * `stream` is a cudaStream_t (NULL = default stream); the call is
 * asynchronous. */
SODA_LAYOUT_API int soda_layout_pack_device(const soda_stream_layout* layout,
                                            const void* dense,
                                            void* const* bank_buffers,
                                            void* stream);

/* Un-tiler: `banks` device buffers -> the valid interior of the dense device
 * array (cells outside it are not touched, like the reference). */
SODA_LAYOUT_API int soda_layout_unpack_device(const soda_stream_layout* layout,
                                              const void* const* bank_buffers,
                                              void* dense, void* stream);

/* Kernels launched by this library since it was loaded. */
SODA_LAYOUT_API int64_t soda_layout_launch_count(void);

SODA_LAYOUT_API const char* soda_layout_last_error(void);

#ifdef __cplusplus
}
#endif

#endif  /* SODA_LAYOUT_H_ */
