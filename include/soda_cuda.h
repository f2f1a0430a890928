/* C ABI of a compiled SODA program for NVIDIA B200 (sm_100a).
 *
 * `sodac prog.soda --cuda-lib libprog.so` builds one shared library per SODA
 * program.  Every such library exports the same symbols (declared here), so a
 * host binds them with dlopen/ctypes/cgo without knowing the program, plus one
 * program-named entry point `soda_cuda_<app>` whose argument list mirrors the
 * host wrapper the reference generates:
 *
 *   reference (src/soda/codegen/frt/host.py:62-89):
 *     int soda::app::<app>(const T* var_<in>_ptr, const int32_t var_<in>_extent[D],
 *                          const int32_t var_<in>_stride[D], const int32_t var_<in>_min[D],
 *                          ... the same four per output (non-const ptr) ...,
 *                          const char* bitstream, int burst_width, int tile_size_<d>...,
 *                          int unroll_factor);
 *   this library:
 *     int soda_cuda_<app>(const T* var_<in>_ptr, const int32_t* var_<in>_extent,
 *                         const int32_t* var_<in>_stride, const int32_t* var_<in>_min,
 *                         ... the same four per output ...,
 *                         const soda_cuda_opts* opts);
 *
 *   `bitstream`, `burst_width`, `tile_size_*`, `unroll_factor` select an FPGA
 *   image and its stream layout; they have no meaning on a GPU and are replaced
 *   by `opts` (may be NULL).  As in the reference, the call blocks, returns 0 on
 *   success, the caller owns every array, only the valid interior of each output
 *   is written (src/soda/codegen/frt/host.py:357-374,422-424) and `min` is
 *   accepted but not used (the reference only logs it, :251-262).
 *
 * Arrays are dense in dimension 0 (stride[0] == 1); stride[d] is the distance
 * in elements between consecutive indices of dimension d.
 *
 * Error convention: 0 = success; otherwise a soda_cuda_status value, and
 * soda_cuda_last_error() returns a message for the calling thread.  No
 * exception crosses this boundary.
 */
#ifndef SODA_CUDA_H_
#define SODA_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Program libraries are built with hidden visibility (and -fno-gnu-unique) so
 * that several of them can live in one process; only this ABI is exported. */
#if defined(__GNUC__)
#define SODA_CUDA_API __attribute__((visibility("default")))
#else
#define SODA_CUDA_API
#endif

#define SODA_CUDA_MAX_DIM 3
#define SODA_CUDA_MAX_TENSORS 8
#define SODA_CUDA_MAX_PASSES 1024

typedef enum soda_cuda_status {
  SODA_CUDA_OK = 0,
  SODA_CUDA_BAD_ARGUMENT = 1,
  SODA_CUDA_CUDA_ERROR = 2,
  SODA_CUDA_UNSUPPORTED = 3,
  SODA_CUDA_OUT_OF_MEMORY = 4,
  SODA_CUDA_COMM_ERROR = 5   /* NCCL or the exchange callback failed */
} soda_cuda_status;

typedef enum soda_cuda_dtype {
  SODA_CUDA_U8 = 0, SODA_CUDA_I8 = 1, SODA_CUDA_U16 = 2, SODA_CUDA_I16 = 3,
  SODA_CUDA_U32 = 4, SODA_CUDA_I32 = 5, SODA_CUDA_U64 = 6, SODA_CUDA_I64 = 7,
  SODA_CUDA_F32 = 8, SODA_CUDA_F64 = 9,
  SODA_CUDA_F16 = 10 /* IEEE binary16, the DSL's `half` */
} soda_cuda_dtype;

/* Run-time options; zero-initialise and set struct_size = sizeof(soda_cuda_opts). */
typedef struct soda_cuda_opts {
  int32_t struct_size;
  int32_t device;        /* CUDA device ordinal; -1 = current device */
  void* stream;          /* cudaStream_t; NULL = the default stream */
  int32_t segment;       /* output slices per CTA along the streamed dimension; 0 = auto:
                          * the first launch of a pass on a large window (>= 2^22 cells)
                          * times a few candidate lengths with CUDA events on `stream`
                          * and synchronises on them once (the pass is idempotent); later
                          * launches of the same shape reuse the winner.  Set a length, or
                          * SODA_CUDA_AUTOTUNE=0 in the environment, when that one-time
                          * synchronisation is unwanted (e.g. while capturing a CUDA graph) */
  int32_t reserved[5];   /* reserved[0]: chunks of the pipelined host path (copy/compute
                          * overlap of soda_cuda_plan_run_host); 0 = auto, 1 = off,
                          * n > 1: n equal chunks, -n: n chunks, each 8 % shorter
                          * than the one before it (the call ends one chunk's passes
                          * and download after the last upload, so the last chunk
                          * should be short; auto does the same with up to 20 chunks)
                          * reserved[1]: soda_cuda_<app> / soda_cuda_run_host split the
                          * grid over this many devices (0 .. n-1) when > 1
                          * (sodac --cuda-gpus; see soda_cuda_multi_run_host) */
} soda_cuda_opts;

/* One pass = one HBM round trip = `time_block` fused iterations. */
typedef struct soda_cuda_pass_info {
  int32_t time_block;
  /* how far the pass output depends on the pass input, per dimension:
   * output cell p reads input cells p+reach_lo .. p+reach_hi (reach_lo <= 0) */
  int32_t reach_lo[SODA_CUDA_MAX_DIM];
  int32_t reach_hi[SODA_CUDA_MAX_DIM];
  int32_t threads_per_cta;
  int32_t smem_bytes;
  int32_t cells_per_lane;
  int32_t strip_cells;      /* cells per warp in dimension 0 */
  int32_t valid_cells[2];   /* stored cells per strip / tile in dims 0 (and 1) */
} soda_cuda_pass_info;

typedef struct soda_cuda_program_info {
  const char* app_name;
  const char* soda_source;     /* the program this library was compiled from */
  int32_t dim;
  int32_t iterate;
  int32_t num_inputs;
  int32_t num_outputs;
  const char* input_names[SODA_CUDA_MAX_TENSORS];
  const char* output_names[SODA_CUDA_MAX_TENSORS];
  int32_t input_dtypes[SODA_CUDA_MAX_TENSORS];   /* soda_cuda_dtype */
  int32_t output_dtypes[SODA_CUDA_MAX_TENSORS];
  /* valid box of output o on a grid of `extent`:
   *   [final_lo[o][d], extent[d] - final_hi[o][d])   per dimension d */
  int32_t final_lo[SODA_CUDA_MAX_TENSORS][SODA_CUDA_MAX_DIM];
  int32_t final_hi[SODA_CUDA_MAX_TENSORS][SODA_CUDA_MAX_DIM];
  int32_t num_passes;
  int32_t strict_fp;            /* 1: compiled with --fmad=false (bit-exact vs g++) */
  int32_t algorithmic_bytes_per_cell_per_pass;
  /* `param` arrays of the program (reference: src/soda/grammar.py ParamStmt;
   * soda::app::<app> takes them after the tensors, frt/host.py:73-79) */
  int32_t num_params;
  const char* param_names[SODA_CUDA_MAX_TENSORS];
  int32_t param_dtypes[SODA_CUDA_MAX_TENSORS];
  int32_t param_elems[SODA_CUDA_MAX_TENSORS];   /* product of the declared sizes */
  /* Dimensions of the program as written.  1 for a 1-D program, which runs as
   * the 2-D program over an N x 1 grid (`dim` == 2: lanes along the only
   * dimension, one slice in the streamed one).  soda_cuda_<app> takes the 1-D
   * quadruples of the source program; the generic entry points take the 2-D
   * extent {N, 1} and strides {1, N}. */
  int32_t source_dim;
} soda_cuda_program_info;

typedef struct soda_cuda_plan soda_cuda_plan;   /* opaque */

/* Program description.  Replaces the `// stencil window size` / kStencilDim
 * constants the reference prints into the host (src/soda/codegen/frt/host.py:686-699). */
SODA_CUDA_API int soda_cuda_info(soda_cuda_program_info* info);
SODA_CUDA_API int soda_cuda_get_pass_info(int32_t pass_index, soda_cuda_pass_info* info);

/* Message for the last non-zero status returned to this thread. */
SODA_CUDA_API const char* soda_cuda_last_error(void);

/* One-shot, host arrays in / host arrays out; host<->device copies included.
 * Generic form of soda_cuda_<app>: per-tensor arrays instead of named arguments.
 * Replaces soda::app::<app> (src/soda/codegen/frt/host.py:62-431). */
SODA_CUDA_API int soda_cuda_run_host(const void* const* in_ptrs, const int32_t* const* in_strides,
                       void* const* out_ptrs, const int32_t* const* out_strides,
                       const int32_t* extent, const soda_cuda_opts* opts);

/* A plan owns the device-side scratch (ping-pong buffers, staging copies) for
 * one grid extent, like the tiled buffers the reference wrapper allocates per
 * call (src/soda/codegen/frt/host.py:149-179), but reusable across calls. */
SODA_CUDA_API int soda_cuda_plan_create(const int32_t* extent, const soda_cuda_opts* opts,
                          soda_cuda_plan** plan);
SODA_CUDA_API int soda_cuda_plan_destroy(soda_cuda_plan* plan);

/* Host arrays through a plan (pinned or pageable memory). */
SODA_CUDA_API int soda_cuda_plan_run_host(soda_cuda_plan* plan,
                            const void* const* in_ptrs, const int32_t* const* in_strides,
                            void* const* out_ptrs, const int32_t* const* out_strides);

/* Device arrays through a plan: all `iterate` iterations, asynchronous on the
 * plan's stream.  pitches[t][0] = elements between rows, pitches[t][1] = between
 * planes (3-D).  Base addresses and row pitches must be 16-byte aligned. */
SODA_CUDA_API int soda_cuda_plan_run_device(soda_cuda_plan* plan,
                              const void* const* d_in, const int64_t (*in_pitches)[2],
                              void* const* d_out, const int64_t (*out_pitches)[2]);

/* One pass on caller-owned device buffers (used by the multi-GPU slab runtime,
 * which exchanges halos between passes).  `extent` is the extent of the local
 * arrays; outputs are written inside [box_lo[o], box_hi[o]) only. */
SODA_CUDA_API int soda_cuda_run_pass(int32_t pass_index, const int32_t* extent,
                       const void* const* d_in, const int64_t (*in_pitches)[2],
                       void* const* d_out, const int64_t (*out_pitches)[2],
                       const int32_t (*box_lo)[SODA_CUDA_MAX_DIM],
                       const int32_t (*box_hi)[SODA_CUDA_MAX_DIM],
                       const soda_cuda_opts* opts);

/* Values of `param` array `index` (dense, row-major over the declared sizes,
 * param_elems[index] elements of host memory).  They are copied into the
 * constant memory of the device `opts->device` (default: the current device)
 * and used by every later launch of this library on that device; set them
 * before launching.  The reference passes params to soda::app::<app> with the
 * same (ptr, extent, stride, min) quadruple as tensors (frt/host.py:73-79);
 * soda_cuda_<app> does too and calls this function. */
SODA_CUDA_API int soda_cuda_set_param(int32_t index, const void* values,
                                      const soda_cuda_opts* opts);

/* ---- multi-GPU: one slab of a larger grid per device ------------------------
 *
 * The reference has no distributed path (its host drives one FPGA image,
 * src/soda/codegen/frt/host.py:62-431); BASELINE.json's north star asks for
 * grids split along the outermost dimension over the GPUs of one box with a
 * halo exchange per time block.  A slab is one rank's share of the global
 * grid: its own slices of the streamed (last) dimension plus ghost slices on
 * both sides.  Passes run in exchange groups - k passes between two halo
 * exchanges, each pass also storing the ghost slices the rest of its group
 * still reads - and the last pass of a group is issued as three launches so
 * that the exchange of the boundary slices overlaps the interior.
 *
 * Two transports move the halos: the library's own NCCL communicator
 * (soda_cuda_slab_opts.nccl_id from soda_cuda_nccl_unique_id, distributed to
 * all ranks by the caller; libnccl.so.2 is loaded at run time), or a callback
 * (soda_cuda_slab_opts.exchange) for hosts that already have a communication
 * layer (torch.distributed, MPI).  One process per GPU, or one process driving
 * several devices with one slab each (soda_cuda_multi_run_host). */

#define SODA_CUDA_NCCL_ID_BYTES 128

typedef struct soda_cuda_slab soda_cuda_slab;   /* opaque */

/* One transfer of a halo exchange: `bytes` bytes at `ptr` (device memory of
 * this slab's device) to or from rank `peer`.  Transfers between two ranks are
 * matched in order, like ncclSend / ncclRecv inside one group. */
typedef struct soda_cuda_halo_op {
  int32_t send;    /* 1: send to `peer`, 0: receive from `peer` */
  int32_t peer;
  void* ptr;
  int64_t bytes;
} soda_cuda_halo_op;

/* Callback transport.  phase 0: start all `num_ops` transfers, ordered after
 * the work already queued on `stream` (cudaStream_t); phase 1 (ops == NULL):
 * make `stream` wait for the transfers started last.  Returns 0 on success. */
typedef int (*soda_cuda_exchange_fn)(void* user, int32_t phase,
                                     const soda_cuda_halo_op* ops, int32_t num_ops,
                                     void* stream);

typedef struct soda_cuda_slab_opts {
  int32_t struct_size;
  int32_t rank;
  int32_t world;
  int32_t device;          /* CUDA device ordinal; -1 = current device */
  void* stream;            /* cudaStream_t of the passes; NULL = default stream */
  int32_t segment;         /* as soda_cuda_opts.segment */
  int32_t exchange_every;  /* passes per halo exchange; 0 = as many as keep the ghost
                            * below 2.5 % of the slab; -1 = all (one exchange per run) */
  int32_t no_overlap;      /* 1: exchange after the whole last pass of a group */
  int32_t host_chunks;     /* chunks of soda_cuda_slab_run_host; 0 = auto */
  soda_cuda_exchange_fn exchange;   /* NULL: the native NCCL transport */
  void* exchange_user;
  const void* nccl_id;     /* SODA_CUDA_NCCL_ID_BYTES bytes, identical on all ranks
                            * (native transport, world > 1) */
  int32_t reserved[8];     /* reserved[0] = 1: dry run - bounds, ghost depths and exchange
                            * groups only (soda_cuda_slab_get_info); nothing is allocated
                            * reserved[1]: where soda_cuda_slab_run_host computes the chunks
                            * that read ghost slices: 0 = by transport (NCCL: in their
                            * natural place, the exchange over NVLink is quick; callback:
                            * last), 1 = natural place, 2 = last */
} soda_cuda_slab_opts;

typedef struct soda_cuda_slab_info {
  int32_t begin, end;              /* global slices this rank owns */
  int32_t local_begin, local_end;  /* global slices its arrays cover (own + ghosts) */
  int32_t ghost_lo, ghost_hi;      /* ghost depth a group needs before its first pass */
  int32_t local_extent[SODA_CUDA_MAX_DIM];
  int64_t pitch[2];                /* elements between rows / planes of every array */
  int32_t num_groups;
  int32_t group_passes[16];        /* passes of the first 16 groups */
} soda_cuda_slab_info;

/* Rank 0 calls this and hands the bytes to every rank (native transport). */
SODA_CUDA_API int soda_cuda_nccl_unique_id(void* id_bytes);

SODA_CUDA_API int soda_cuda_slab_create(const int32_t* global_extent,
                                        const soda_cuda_slab_opts* opts,
                                        soda_cuda_slab** slab);
SODA_CUDA_API int soda_cuda_slab_destroy(soda_cuda_slab* slab);
SODA_CUDA_API int soda_cuda_slab_get_info(const soda_cuda_slab* slab,
                                          soda_cuda_slab_info* info);

/* Device addresses of the slab's local arrays (num_inputs + num_outputs
 * pointers; shape local_extent, pitches `pitch`).  The caller fills its own
 * slices [begin - local_begin, end - local_begin) of the inputs; ghost slices
 * are refreshed by soda_cuda_slab_run. */
SODA_CUDA_API int soda_cuda_slab_buffers(soda_cuda_slab* slab, void** d_in, void** d_out);

/* All `iterate` iterations, device-resident: inputs -> outputs.  Queued on the
 * slab's stream; returns after everything (passes and exchanges) is queued. */
SODA_CUDA_API int soda_cuda_slab_run(soda_cuda_slab* slab);

/* One halo exchange of the input arrays, `depth_lo` / `depth_hi` ghost slices
 * (benchmarks time this on its own). */
SODA_CUDA_API int soda_cuda_slab_exchange_inputs(soda_cuda_slab* slab, int32_t depth_lo,
                                                 int32_t depth_hi);

/* Host arrays holding this rank's own slices (extent = the global extent with
 * the last dimension replaced by end - begin) in, the same slices of the
 * outputs out: chunked H2D / passes / D2H pipeline per rank, the ghost slices
 * of the inputs arrive from the neighbouring ranks' uploads while the interior
 * chunks compute.  Needs a slab created with exchange_every = -1.  Blocks. */
SODA_CUDA_API int soda_cuda_slab_run_host(soda_cuda_slab* slab,
                            const void* const* in_ptrs, const int32_t* const* in_strides,
                            void* const* out_ptrs, const int32_t* const* out_strides);

/* One process, `num_devices` GPUs (sodac --cuda-gpus N): the whole grid in host
 * arrays, as soda_cuda_run_host; every device uploads its slab and the ghost
 * slices it needs straight from the host arrays, so the devices never talk to
 * each other.  `devices` may be NULL (0 .. num_devices-1).  Blocks. */
SODA_CUDA_API int soda_cuda_multi_run_host(const void* const* in_ptrs,
                             const int32_t* const* in_strides,
                             void* const* out_ptrs, const int32_t* const* out_strides,
                             const int32_t* extent, const int32_t* devices,
                             int32_t num_devices, const soda_cuda_opts* opts);

/* Page-locked host memory on the NUMA node of `device` (first-touch under an
 * mbind policy when the box has several nodes, then cudaHostRegister), for the
 * staging arrays of the host entry points. */
SODA_CUDA_API int soda_cuda_host_alloc(void** ptr, int64_t bytes, int32_t device);
SODA_CUDA_API int soda_cuda_host_free(void* ptr, int64_t bytes);

/* Number of kernel launches issued by this library since it was loaded. */
SODA_CUDA_API int64_t soda_cuda_launch_count(void);

#ifdef __cplusplus
}  /* extern "C" */
#endif

#endif  /* SODA_CUDA_H_ */
