// Multi-GPU slab runtime behind the C ABI (include/soda_cuda.h, "multi-GPU"
// section).  Included at the end of soda_runtime.cuh.
//
// The grid is split along the outermost (streamed) dimension - the dimension
// the reference itself treats as unbounded (reference: README.md:223; its host
// tiles every other dimension, src/soda/codegen/frt/host.py:124-131) - into
// contiguous slabs, one per rank.  Every slab's arrays cover its own slices
// plus the ghost slices that exist in the global grid; at the global border
// there is no ghost and the kernels' TMA loads zero-fill exactly as on one GPU.
// Every stored value therefore has the same dependency cone and the same
// operation order as in the single-GPU run: results are bit-identical.
//
// Exchange groups are temporal blocking one level up: a rank that holds
// k x reach ghost slices runs k passes without talking to anybody (pass j also
// stores the ghost slices the rest of the group still reads), then swaps
// k x reach slices at once.
#pragma once

#ifndef SODA_EMU
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif

namespace soda {
namespace rt {

// ---- transports -----------------------------------------------------------------
struct Transport {
  virtual ~Transport() {}
  // starts the transfers, ordered after what is queued on `after`
  virtual int start(const std::vector<soda_cuda_halo_op>& ops, cudaStream_t after) = 0;
  // makes `stream` wait for the transfers started last
  virtual int wait(cudaStream_t stream) = 0;
};

struct CallbackTransport : Transport {
  soda_cuda_exchange_fn fn;
  void* user;
  CallbackTransport(soda_cuda_exchange_fn f, void* u) : fn(f), user(u) {}
  int start(const std::vector<soda_cuda_halo_op>& ops, cudaStream_t after) override {
    if (ops.empty()) return SODA_CUDA_OK;
    if (fn(user, 0, ops.data(), static_cast<int32_t>(ops.size()), after) != 0)
      return fail(SODA_CUDA_COMM_ERROR, "the exchange callback failed (start)");
    return SODA_CUDA_OK;
  }
  int wait(cudaStream_t stream) override {
    if (fn(user, 1, nullptr, 0, stream) != 0)
      return fail(SODA_CUDA_COMM_ERROR, "the exchange callback failed (wait)");
    return SODA_CUDA_OK;
  }
};

#ifndef SODA_EMU
// The slice of NCCL this runtime uses, bound at run time: a program library
// has no link-time dependency on NCCL, and inside a process that already
// loaded a libnccl.so.2 (PyTorch) the same copy is used.
struct NcclApi {
  struct UniqueId { char bytes[SODA_CUDA_NCCL_ID_BYTES]; };
  typedef void* Comm;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;

  static NcclApi& get() {
    static NcclApi api = [] {
      NcclApi a;
      void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (lib == nullptr) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (lib == nullptr) {
        a.why = std::string("cannot load libnccl.so.2: ") + dlerror();
        return a;
      }
      bool all = true;
      auto bind = [&](const char* name, void* target) {
        void* sym = dlsym(lib, name);
        if (sym == nullptr) {
          all = false;
          a.why = std::string("libnccl lacks ") + name;
        }
        memcpy(target, &sym, sizeof(sym));
      };
      bind("ncclGetUniqueId", &a.GetUniqueId);
      bind("ncclCommInitRank", &a.CommInitRank);
      bind("ncclCommDestroy", &a.CommDestroy);
      bind("ncclGroupStart", &a.GroupStart);
      bind("ncclGroupEnd", &a.GroupEnd);
      bind("ncclSend", &a.Send);
      bind("ncclRecv", &a.Recv);
      bind("ncclGetErrorString", &a.GetErrorString);
      a.ok = all;
      return a;
    }();
    return api;
  }
  int check(int result, const char* what) const {
    if (result == 0) return SODA_CUDA_OK;
    return fail(SODA_CUDA_COMM_ERROR, std::string(what) + ": " +
                                          (GetErrorString ? GetErrorString(result)
                                                          : "NCCL error"));
  }
};

struct NcclTransport : Transport {
  NcclApi::Comm comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ready = nullptr, done = nullptr;

  int init(const void* id_bytes, int rank, int world) {
    NcclApi& api = NcclApi::get();
    if (!api.ok) return fail(SODA_CUDA_COMM_ERROR, api.why);
    NcclApi::UniqueId id;
    memcpy(id.bytes, id_bytes, sizeof(id.bytes));
    int status = api.check(api.CommInitRank(&comm, world, id, rank),
                           "ncclCommInitRank");
    if (status != SODA_CUDA_OK) return status;
    SODA_CUDA_CHECK(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
    SODA_CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    SODA_CUDA_CHECK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    return SODA_CUDA_OK;
  }
  ~NcclTransport() override {
    if (comm) NcclApi::get().CommDestroy(comm);
    if (ready) cudaEventDestroy(ready);
    if (done) cudaEventDestroy(done);
    if (comm_stream) cudaStreamDestroy(comm_stream);
  }
  int start(const std::vector<soda_cuda_halo_op>& ops, cudaStream_t after) override {
    NcclApi& api = NcclApi::get();
    SODA_CUDA_CHECK(cudaEventRecord(ready, after));
    SODA_CUDA_CHECK(cudaStreamWaitEvent(comm_stream, ready, 0));
    if (!ops.empty()) {
      int status = api.check(api.GroupStart(), "ncclGroupStart");
      for (size_t i = 0; i < ops.size() && status == SODA_CUDA_OK; ++i) {
        const soda_cuda_halo_op& op = ops[i];
        const size_t bytes = static_cast<size_t>(op.bytes);
        status = op.send
                     ? api.check(api.Send(op.ptr, bytes, /*ncclInt8*/ 0, op.peer,
                                          comm, comm_stream), "ncclSend")
                     : api.check(api.Recv(op.ptr, bytes, /*ncclInt8*/ 0, op.peer,
                                          comm, comm_stream), "ncclRecv");
      }
      int end_status = api.check(api.GroupEnd(), "ncclGroupEnd");
      if (status != SODA_CUDA_OK) return status;
      if (end_status != SODA_CUDA_OK) return end_status;
    }
    SODA_CUDA_CHECK(cudaEventRecord(done, comm_stream));
    return SODA_CUDA_OK;
  }
  int wait(cudaStream_t stream) override {
    SODA_CUDA_CHECK(cudaStreamWaitEvent(stream, done, 0));
    return SODA_CUDA_OK;
  }
};
#endif  // SODA_EMU

inline void split_slices(int total, int world, std::vector<int>* begin) {
  begin->assign(world + 1, 0);
  const int base = total / world, rest = total % world;
  for (int r = 0; r < world; ++r)
    (*begin)[r + 1] = (*begin)[r] + base + (r < rest ? 1 : 0);
}

}  // namespace rt
}  // namespace soda

struct soda_cuda_slab {
  int rank = 0, world = 1;
  int global_extent[soda::rt::kMaxD] = {0, 0, 0};
  int begin = 0, end = 0;              // global slices owned
  int local_begin = 0, local_end = 0;  // global slices the arrays cover
  int own_lo = 0, own_hi = 0;          // the same, local
  int ghost_lo = 0, ghost_hi = 0;
  bool no_overlap = false;
  bool dry_run = false;  // described, not allocated: only get_info is valid
  std::vector<std::pair<int, int>> pass_reach;  // (lo, hi) >= 0 per pass
  std::vector<std::vector<int>> groups;
  soda_cuda_plan* plan = nullptr;  // local arrays, scratch, streams
  soda::rt::Transport* transport = nullptr;
  bool defer_edge_chunks = true;  // host pipeline: see HostPipeline
};

namespace soda {
namespace rt {

inline void slab_make_groups(soda_cuda_slab* slab, int exchange_every, int min_slab) {
  const int n = static_cast<int>(slab->pass_reach.size());
  slab->groups.clear();
  if (slab->world == 1 || exchange_every < 0) {
    std::vector<int> all(n);
    for (int i = 0; i < n; ++i) all[i] = i;
    slab->groups.push_back(all);
    return;
  }
  // default: as many passes as keep the ghost below 2.5 % of the slab
  const int budget = std::max(1, static_cast<int>(min_slab * 0.025));
  std::vector<int> current;
  int depth = 0;
  for (int i = 0; i < n; ++i) {
    const int reach = std::max(slab->pass_reach[i].first, slab->pass_reach[i].second);
    const bool full = exchange_every > 0
                          ? static_cast<int>(current.size()) >= exchange_every
                          : depth + reach > budget;
    if (!current.empty() && full) {
      slab->groups.push_back(current);
      current.clear();
      depth = 0;
    }
    current.push_back(i);
    depth += reach;
  }
  slab->groups.push_back(current);
}

inline void slab_group_depth(const soda_cuda_slab* slab, const std::vector<int>& group,
                             size_t from, int* lo, int* hi) {
  *lo = *hi = 0;
  for (size_t k = from; k < group.size(); ++k) {
    *lo += slab->pass_reach[group[k]].first;
    *hi += slab->pass_reach[group[k]].second;
  }
}

// The transfers that refresh `depth_lo` / `depth_hi` ghost slices of `arrays`:
// a rank's lower ghost comes from the top of the rank below, its upper ghost
// from the bottom of the rank above.
inline void slab_exchange_ops(const soda_cuda_slab* slab, void* const* arrays,
                              const int* elem_bytes, int count, int depth_lo,
                              int depth_hi, std::vector<soda_cuda_halo_op>* ops) {
  const ProgramDesc& prog = soda_program();
  const soda_cuda_plan* plan = slab->plan;
  const long long slice_cells = prog.info.dim == 2 ? plan->pitch[0] : plan->pitch[1];
  auto add = [&](int send, int peer, void* base, int elem, int first, int slices) {
    soda_cuda_halo_op op;
    op.send = send;
    op.peer = peer;
    op.ptr = static_cast<char*>(base) + slice_cells * first * elem;
    op.bytes = slice_cells * slices * elem;
    ops->push_back(op);
  };
  for (int t = 0; t < count; ++t) {
    if (slab->rank > 0) {
      if (depth_hi > 0)  // the lower neighbour's upper ghost is my bottom slices
        add(1, slab->rank - 1, arrays[t], elem_bytes[t], slab->own_lo, depth_hi);
      if (depth_lo > 0)
        add(0, slab->rank - 1, arrays[t], elem_bytes[t], slab->own_lo - depth_lo,
            depth_lo);
    }
    if (slab->rank < slab->world - 1) {
      if (depth_lo > 0)  // the upper neighbour's lower ghost is my top slices
        add(1, slab->rank + 1, arrays[t], elem_bytes[t], slab->own_hi - depth_lo,
            depth_lo);
      if (depth_hi > 0)
        add(0, slab->rank + 1, arrays[t], elem_bytes[t], slab->own_hi, depth_hi);
    }
  }
}

inline int slab_start_exchange(soda_cuda_slab* slab, void* const* arrays,
                               const int* elem_bytes, int count, int depth_lo,
                               int depth_hi, cudaStream_t after) {
  if (slab->world == 1) return SODA_CUDA_OK;
  std::vector<soda_cuda_halo_op> ops;
  slab_exchange_ops(slab, arrays, elem_bytes, count, depth_lo, depth_hi, &ops);
  return slab->transport->start(ops, after);
}

inline int slab_wait_exchange(soda_cuda_slab* slab, cudaStream_t stream) {
  if (slab->world == 1) return SODA_CUDA_OK;
  return slab->transport->wait(stream);
}

// One launch of pass `index`: `current` -> `target`, outputs stored in the
// local slices [lo_s, hi_s) (clipped to the final valid box on the last pass).
inline int slab_launch(soda_cuda_slab* slab, int index, void* const* current,
                       void* const* target, int lo_s, int hi_s) {
  const ProgramDesc& prog = soda_program();
  soda_cuda_plan* plan = slab->plan;
  const int dim = prog.info.dim, s_dim = dim - 1;
  const bool last = index == prog.info.num_passes - 1;
  PassArgs a;
  memset(&a, 0, sizeof(a));
  for (int d = 0; d < dim; ++d) a.extent[d] = plan->extent[d];
  a.segment = plan->segment;
  a.stream = plan->stream;
  for (int i = 0; i < prog.info.num_inputs; ++i) {
    a.in[i] = current[i];
    a.in_pitch[i][0] = plan->pitch[0];
    a.in_pitch[i][1] = plan->pitch[1];
  }
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    a.out[o] = target[o];
    a.out_pitch[o][0] = plan->pitch[0];
    a.out_pitch[o][1] = plan->pitch[1];
  }
  default_boxes(prog, plan, last, a.box_lo, a.box_hi);
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    const int lo = std::max(a.box_lo[o][s_dim], lo_s);
    const int hi = std::min(a.box_hi[o][s_dim], hi_s);
    a.box_lo[o][s_dim] = lo;
    a.box_hi[o][s_dim] = std::max(lo, hi);
  }
  return launch_tuned(prog, prog.schedule[index], a);
}

inline int slab_run(soda_cuda_slab* slab) {
  const ProgramDesc& prog = soda_program();
  soda_cuda_plan* plan = slab->plan;
  const int dim = prog.info.dim, s_dim = dim - 1;
  const int n_out = prog.info.num_outputs;
  const int total = slab->global_extent[s_dim];
  const int local = plan->extent[s_dim];
  void* const* current = plan->d_in;
  const int* current_bytes = prog.in_elem_bytes;

  int depth_lo = 0, depth_hi = 0;
  slab_group_depth(slab, slab->groups[0], 0, &depth_lo, &depth_hi);
  int status = slab_start_exchange(slab, current, current_bytes,
                                   prog.info.num_inputs, depth_lo, depth_hi,
                                   plan->stream);
  if (status == SODA_CUDA_OK) status = slab_wait_exchange(slab, plan->stream);
  if (status != SODA_CUDA_OK) return status;

  for (size_t g = 0; g < slab->groups.size(); ++g) {
    const std::vector<int>& group = slab->groups[g];
    for (size_t k = 0; k < group.size(); ++k) {
      const int index = group[k];
      const bool last = index == prog.info.num_passes - 1;
      void* const* target;
      if (last) {
        target = plan->d_out;
      } else {
        for (int o = 0; o < n_out; ++o) {
          status = plan_alloc(plan, &plan->scratch[index & 1][o],
                              prog.out_elem_bytes[o]);
          if (status != SODA_CUDA_OK) return status;
        }
        target = plan->scratch[index & 1];
      }
      // ghost slices the rest of the group still needs from this pass
      int rest_lo = 0, rest_hi = 0;
      slab_group_depth(slab, group, k + 1, &rest_lo, &rest_hi);
      const int lo_slice = slab->begin > 0 ? std::max(0, slab->own_lo - rest_lo)
                                           : slab->own_lo;
      const int hi_slice = slab->end < total ? std::min(local, slab->own_hi + rest_hi)
                                             : slab->own_hi;
      const int store_lo = last ? slab->own_lo : lo_slice;
      const int store_hi = last ? slab->own_hi : std::max(lo_slice, hi_slice);
      const bool end_of_group = k + 1 == group.size();
      if (last || slab->world == 1 || !end_of_group) {
        status = slab_launch(slab, index, current, target, store_lo, store_hi);
        if (status != SODA_CUDA_OK) return status;
        current = target;
        current_bytes = prog.out_elem_bytes;
        continue;
      }
      int next_lo = 0, next_hi = 0;
      slab_group_depth(slab, slab->groups[g + 1], 0, &next_lo, &next_hi);
      // slices of this pass's output that a neighbour needs for the next
      // group: computed first, sent while the interior runs
      int bottom_lo = slab->own_lo, bottom_hi = slab->own_lo;
      if (slab->rank > 0) bottom_hi = std::min(slab->own_hi, slab->own_lo + next_hi);
      int top_lo = slab->own_hi, top_hi = slab->own_hi;
      if (slab->rank < slab->world - 1)
        top_lo = std::max(bottom_hi, slab->own_hi - next_lo);
      if (slab->no_overlap) {
        bottom_hi = bottom_lo;
        top_lo = top_hi;
      }
      if (bottom_hi > bottom_lo)
        status = slab_launch(slab, index, current, target,
                             std::max(store_lo, bottom_lo),
                             std::min(store_hi, bottom_hi));
      if (status == SODA_CUDA_OK && top_hi > top_lo)
        status = slab_launch(slab, index, current, target,
                             std::max(store_lo, top_lo), std::min(store_hi, top_hi));
      if (status == SODA_CUDA_OK)
        status = slab_start_exchange(slab, target, prog.out_elem_bytes, n_out,
                                     next_lo, next_hi, plan->stream);
      if (status == SODA_CUDA_OK && top_lo > bottom_hi)
        status = slab_launch(slab, index, current, target,
                             std::max(store_lo, bottom_hi),
                             std::min(store_hi, top_lo));
      if (status == SODA_CUDA_OK) status = slab_wait_exchange(slab, plan->stream);
      if (status != SODA_CUDA_OK) return status;
      current = target;
      current_bytes = prog.out_elem_bytes;
    }
  }
  return SODA_CUDA_OK;
}

// A plan over the slices [local_begin, local_end) of a larger grid.
inline int slab_make_plan(const int32_t* global_extent, int local_begin,
                          int local_end, int device, cudaStream_t stream,
                          int segment, int host_chunks, soda_cuda_plan** out) {
  const ProgramDesc& prog = soda_program();
  const int dim = prog.info.dim, s_dim = dim - 1;
  int32_t local_extent[kMaxD] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) local_extent[d] = global_extent[d];
  local_extent[s_dim] = local_end - local_begin;
  soda_cuda_opts opts;
  memset(&opts, 0, sizeof(opts));
  opts.struct_size = sizeof(opts);
  opts.device = device;
  opts.stream = stream;
  opts.segment = segment;
  opts.reserved[0] = host_chunks;
  int status = soda_cuda_plan_create(local_extent, &opts, out);
  if (status != SODA_CUDA_OK) return status;
  for (int o = 0; o < prog.info.num_outputs; ++o) {
    // the valid range of the global grid, in local coordinates
    (*out)->s_valid_lo[o] = prog.info.final_lo[o][s_dim] - local_begin;
    (*out)->s_valid_hi[o] =
        global_extent[s_dim] - prog.info.final_hi[o][s_dim] - local_begin;
  }
  return SODA_CUDA_OK;
}

struct SlabHooks {
  soda_cuda_slab* slab;
  int depth_lo, depth_hi;
  static int after_edges(void* user, cudaStream_t copy_in) {
    SlabHooks* h = static_cast<SlabHooks*>(user);
    const ProgramDesc& prog = soda_program();
    return slab_start_exchange(h->slab, h->slab->plan->d_in, prog.in_elem_bytes,
                               prog.info.num_inputs, h->depth_lo, h->depth_hi,
                               copy_in);
  }
  static int before_edge(void* user, cudaStream_t compute) {
    return slab_wait_exchange(static_cast<SlabHooks*>(user)->slab, compute);
  }
};

}  // namespace rt
}  // namespace soda

extern "C" {

int soda_cuda_nccl_unique_id(void* id_bytes) {
  using namespace soda::rt;
  if (id_bytes == nullptr) return fail(SODA_CUDA_BAD_ARGUMENT, "id is NULL");
#ifdef SODA_EMU
  return fail(SODA_CUDA_UNSUPPORTED, "no NCCL in the emulation");
#else
  NcclApi& api = NcclApi::get();
  if (!api.ok) return fail(SODA_CUDA_COMM_ERROR, api.why);
  NcclApi::UniqueId id;
  int status = api.check(api.GetUniqueId(&id), "ncclGetUniqueId");
  if (status != SODA_CUDA_OK) return status;
  memcpy(id_bytes, id.bytes, sizeof(id.bytes));
  return SODA_CUDA_OK;
#endif
}

int soda_cuda_slab_destroy(soda_cuda_slab* slab) {
  if (slab == nullptr) return SODA_CUDA_OK;
  if (slab->dry_run) {
    delete slab->plan;
    delete slab;
    return SODA_CUDA_OK;
  }
  soda::rt::DeviceGuard guard;
  if (slab->plan) guard.enter(slab->plan->device);
  delete slab->transport;
  soda_cuda_plan_destroy(slab->plan);
  delete slab;
  return SODA_CUDA_OK;
}

int soda_cuda_slab_create(const int32_t* global_extent, const soda_cuda_slab_opts* opts,
                          soda_cuda_slab** out) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (global_extent == nullptr || opts == nullptr || out == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "extent / opts / slab is NULL");
  if (opts->world < 1 || opts->rank < 0 || opts->rank >= opts->world)
    return fail(SODA_CUDA_BAD_ARGUMENT, "bad rank / world");
  const int dim = prog.info.dim, s_dim = dim - 1;
  for (int d = 0; d < dim; ++d)
    if (global_extent[d] <= 0)
      return fail(SODA_CUDA_BAD_ARGUMENT, "extent must be positive");
  if (prog.info.num_passes > 1 && prog.info.num_inputs != prog.info.num_outputs)
    return fail(SODA_CUDA_UNSUPPORTED, "iterated programs map outputs to inputs");
  soda_cuda_slab* slab = new soda_cuda_slab();
  slab->rank = opts->rank;
  slab->world = opts->world;
  slab->no_overlap = opts->no_overlap != 0;
  for (int d = 0; d < dim; ++d) slab->global_extent[d] = global_extent[d];
  const int total = global_extent[s_dim];
  std::vector<int> bounds;
  split_slices(total, opts->world, &bounds);
  slab->begin = bounds[opts->rank];
  slab->end = bounds[opts->rank + 1];
  for (int pass = 0; pass < prog.info.num_passes; ++pass) {
    const soda_cuda_pass_info& info = prog.impls[prog.schedule[pass]].info;
    slab->pass_reach.emplace_back(std::max(0, -info.reach_lo[s_dim]),
                                  std::max(0, info.reach_hi[s_dim]));
  }
  int min_slab = total;
  for (int r = 0; r < opts->world; ++r)
    min_slab = std::min(min_slab, bounds[r + 1] - bounds[r]);
  slab_make_groups(slab, opts->exchange_every, min_slab);
  for (const std::vector<int>& group : slab->groups) {
    int lo = 0, hi = 0;
    slab_group_depth(slab, group, 0, &lo, &hi);
    slab->ghost_lo = std::max(slab->ghost_lo, lo);
    slab->ghost_hi = std::max(slab->ghost_hi, hi);
  }
  if (opts->world > 1 && min_slab < std::max(slab->ghost_lo, slab->ghost_hi)) {
    delete slab;
    return fail(SODA_CUDA_BAD_ARGUMENT,
                "slab thinner than the halo: use fewer ranks or exchange more often");
  }
  slab->local_begin = std::max(0, slab->begin - slab->ghost_lo);
  slab->local_end = std::min(total, slab->end + slab->ghost_hi);
  slab->own_lo = slab->begin - slab->local_begin;
  slab->own_hi = slab->end - slab->local_begin;

  if (opts->reserved[0] == 1) {
    // dry run (planning tools, tests): bounds, ghosts and groups only
    slab->plan = new soda_cuda_plan();
    memset(slab->plan, 0, sizeof(*slab->plan));
    for (int d = 0; d < dim; ++d) slab->plan->extent[d] = global_extent[d];
    slab->plan->extent[s_dim] = slab->local_end - slab->local_begin;
    slab->plan->device = -1;
    slab->plan->pitch[0] = (global_extent[0] + 127) / 128 * 128;
    slab->plan->pitch[1] = dim == 3 ? slab->plan->pitch[0] * global_extent[1] : 0;
    slab->dry_run = true;
    *out = slab;
    return SODA_CUDA_OK;
  }
  DeviceGuard guard;
  int status = guard.enter(opts->device);
  if (status == SODA_CUDA_OK)
    status = slab_make_plan(global_extent, slab->local_begin, slab->local_end,
                            opts->device, static_cast<cudaStream_t>(opts->stream),
                            opts->segment, opts->host_chunks, &slab->plan);
  soda_cuda_plan* plan = slab->plan;
  for (int i = 0; status == SODA_CUDA_OK && i < prog.info.num_inputs; ++i) {
    status = plan_alloc(plan, &plan->d_in[i], prog.in_elem_bytes[i]);
    if (status == SODA_CUDA_OK &&
        cudaMemsetAsync(plan->d_in[i], 0, plan_cells(plan, dim) * prog.in_elem_bytes[i],
                        plan->stream) != cudaSuccess)
      status = fail(SODA_CUDA_CUDA_ERROR, "cudaMemsetAsync failed");
  }
  for (int o = 0; status == SODA_CUDA_OK && o < prog.info.num_outputs; ++o) {
    status = plan_alloc(plan, &plan->d_out[o], prog.out_elem_bytes[o]);
    if (status == SODA_CUDA_OK &&
        cudaMemsetAsync(plan->d_out[o], 0,
                        plan_cells(plan, dim) * prog.out_elem_bytes[o],
                        plan->stream) != cudaSuccess)
      status = fail(SODA_CUDA_CUDA_ERROR, "cudaMemsetAsync failed");
  }
  if (status == SODA_CUDA_OK && opts->world > 1) {
    if (opts->exchange != nullptr) {
      slab->transport = new CallbackTransport(opts->exchange, opts->exchange_user);
    } else {
#ifdef SODA_EMU
      status = fail(SODA_CUDA_UNSUPPORTED, "the emulation has no NCCL: pass a callback");
#else
      if (opts->nccl_id == nullptr) {
        status = fail(SODA_CUDA_BAD_ARGUMENT,
                      "world > 1 needs opts->nccl_id (soda_cuda_nccl_unique_id) or "
                      "an exchange callback");
      } else {
        NcclTransport* nccl = new NcclTransport();
        slab->transport = nccl;
        slab->defer_edge_chunks = false;  // NVLink: the exchange is quick
        status = nccl->init(opts->nccl_id, opts->rank, opts->world);
      }
#endif
    }
  }
  if (opts->reserved[1] == 1) slab->defer_edge_chunks = false;
  if (opts->reserved[1] == 2) slab->defer_edge_chunks = true;
  if (status != SODA_CUDA_OK) {
    soda_cuda_slab_destroy(slab);
    return status;
  }
  *out = slab;
  return SODA_CUDA_OK;
}

int soda_cuda_slab_get_info(const soda_cuda_slab* slab, soda_cuda_slab_info* info) {
  using namespace soda::rt;
  if (slab == nullptr || info == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "slab / info is NULL");
  memset(info, 0, sizeof(*info));
  info->begin = slab->begin;
  info->end = slab->end;
  info->local_begin = slab->local_begin;
  info->local_end = slab->local_end;
  info->ghost_lo = slab->ghost_lo;
  info->ghost_hi = slab->ghost_hi;
  for (int d = 0; d < kMaxD; ++d) info->local_extent[d] = slab->plan->extent[d];
  info->pitch[0] = slab->plan->pitch[0];
  info->pitch[1] = slab->plan->pitch[1];
  info->num_groups = static_cast<int32_t>(slab->groups.size());
  for (size_t g = 0; g < slab->groups.size() && g < 16; ++g)
    info->group_passes[g] = static_cast<int32_t>(slab->groups[g].size());
  return SODA_CUDA_OK;
}

int soda_cuda_slab_buffers(soda_cuda_slab* slab, void** d_in, void** d_out) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (slab == nullptr || d_in == nullptr || d_out == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "NULL argument");
  if (slab->dry_run) return fail(SODA_CUDA_BAD_ARGUMENT, "slab was created as a dry run");
  for (int i = 0; i < prog.info.num_inputs; ++i) d_in[i] = slab->plan->d_in[i];
  for (int o = 0; o < prog.info.num_outputs; ++o) d_out[o] = slab->plan->d_out[o];
  return SODA_CUDA_OK;
}

int soda_cuda_slab_run(soda_cuda_slab* slab) {
  using namespace soda::rt;
  if (slab == nullptr) return fail(SODA_CUDA_BAD_ARGUMENT, "slab is NULL");
  if (slab->dry_run) return fail(SODA_CUDA_BAD_ARGUMENT, "slab was created as a dry run");
  DeviceGuard guard;
  int status = guard.enter(slab->plan->device);
  if (status != SODA_CUDA_OK) return status;
  return slab_run(slab);
}

int soda_cuda_slab_exchange_inputs(soda_cuda_slab* slab, int32_t depth_lo,
                                   int32_t depth_hi) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (slab == nullptr || slab->dry_run)
    return fail(SODA_CUDA_BAD_ARGUMENT, "slab is NULL or a dry run");
  // a side without a neighbour has no ghost and takes no part in the exchange
  const bool below = slab->rank > 0, above = slab->rank < slab->world - 1;
  if (depth_lo < 0 || depth_hi < 0 || (below && depth_lo > slab->own_lo) ||
      (above &&
       depth_hi > slab->plan->extent[prog.info.dim - 1] - slab->own_hi) ||
      depth_lo > slab->own_hi - slab->own_lo ||
      depth_hi > slab->own_hi - slab->own_lo)
    return fail(SODA_CUDA_BAD_ARGUMENT, "exchange deeper than the ghost");
  DeviceGuard guard;
  int status = guard.enter(slab->plan->device);
  if (status != SODA_CUDA_OK) return status;
  status = slab_start_exchange(slab, slab->plan->d_in, prog.in_elem_bytes,
                               prog.info.num_inputs, depth_lo, depth_hi,
                               slab->plan->stream);
  if (status != SODA_CUDA_OK) return status;
  return slab_wait_exchange(slab, slab->plan->stream);
}

int soda_cuda_slab_run_host(soda_cuda_slab* slab, const void* const* in_ptrs,
                            const int32_t* const* in_strides, void* const* out_ptrs,
                            const int32_t* const* out_strides) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (slab == nullptr || in_ptrs == nullptr || out_ptrs == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "NULL argument");
  if (slab->dry_run) return fail(SODA_CUDA_BAD_ARGUMENT, "slab was created as a dry run");
  if (slab->groups.size() != 1)
    return fail(SODA_CUDA_UNSUPPORTED,
                "soda_cuda_slab_run_host needs one exchange group: create the slab "
                "with exchange_every = -1");
  soda_cuda_plan* plan = slab->plan;
  const int dim = prog.info.dim;
  DeviceGuard guard;
  int status = guard.enter(plan->device);
  if (status != SODA_CUDA_OK) return status;
  int host_extent[kMaxD] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) host_extent[d] = plan->extent[d];
  host_extent[dim - 1] = slab->end - slab->begin;
  status = soda_plan_prepare_host(plan, in_ptrs, in_strides, out_ptrs, out_strides,
                                  host_extent);
  if (status != SODA_CUDA_OK) return status;
  SlabHooks hooks{slab, slab->ghost_lo, slab->ghost_hi};
  HostPipeline pipe;
  pipe.plan = plan;
  pipe.in_ptrs = in_ptrs;
  pipe.in_strides = in_strides;
  pipe.out_ptrs = out_ptrs;
  pipe.out_strides = out_strides;
  pipe.own_lo = pipe.host_lo = slab->own_lo;
  pipe.own_hi = pipe.host_hi = slab->own_hi;
  pipe.host_shift = -slab->own_lo;
  pipe.ghosts_from_peers = slab->world > 1;
  pipe.defer_edge_chunks = slab->defer_edge_chunks;
  pipe.after_edges = &SlabHooks::after_edges;
  pipe.before_edge = &SlabHooks::before_edge;
  pipe.user = &hooks;
  return pipe.finish(pipe.issue());
}

int soda_cuda_multi_run_host(const void* const* in_ptrs, const int32_t* const* in_strides,
                             void* const* out_ptrs, const int32_t* const* out_strides,
                             const int32_t* extent, const int32_t* devices,
                             int32_t num_devices, const soda_cuda_opts* opts) {
  using namespace soda::rt;
  const ProgramDesc& prog = soda_program();
  if (in_ptrs == nullptr || out_ptrs == nullptr || extent == nullptr)
    return fail(SODA_CUDA_BAD_ARGUMENT, "NULL argument");
  const int dim = prog.info.dim, s_dim = dim - 1;
  int available = 0;
  SODA_CUDA_CHECK(cudaGetDeviceCount(&available));
  if (available == 0)
    return fail(SODA_CUDA_CUDA_ERROR, "no CUDA device: this library has no CPU path");
  if (num_devices < 1 || num_devices > available)
    return fail(SODA_CUDA_BAD_ARGUMENT, "bad number of devices");
  for (int d = 0; d < dim; ++d)
    if (extent[d] <= 0) return fail(SODA_CUDA_BAD_ARGUMENT, "extent must be positive");
  const int total = extent[s_dim];
  int reach_lo = 0, reach_hi = 0;
  total_reach(prog, &reach_lo, &reach_hi);
  // slabs thinner than the halo gain nothing: use fewer devices
  int world = num_devices;
  while (world > 1 && total / world < std::max(1, std::max(reach_lo, reach_hi))) --world;
  std::vector<int> bounds;
  split_slices(total, world, &bounds);
  int host_extent[kMaxD] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) host_extent[d] = extent[d];

  std::vector<soda_cuda_plan*> plans(world, nullptr);
  std::vector<HostPipeline> pipes(world);
  std::vector<int> issued(world, SODA_CUDA_OK);
  int status = SODA_CUDA_OK;
  int previous = 0;
  cudaGetDevice(&previous);
  for (int r = 0; r < world && status == SODA_CUDA_OK; ++r) {
    const int device = devices ? devices[r] : r;
    const int local_begin = std::max(0, bounds[r] - reach_lo);
    const int local_end = std::min(total, bounds[r + 1] + reach_hi);
    if (cudaSetDevice(device) != cudaSuccess) {
      status = fail(SODA_CUDA_CUDA_ERROR, "cudaSetDevice failed");
      break;
    }
    // one stream per device: the caller's stream belongs to one device only
    status = slab_make_plan(extent, local_begin, local_end, device, nullptr,
                            opts ? opts->segment : 0, opts ? opts->reserved[0] : 0,
                            &plans[r]);
    if (status != SODA_CUDA_OK) break;
    if (cudaStreamCreateWithFlags(&plans[r]->stream, cudaStreamNonBlocking) !=
        cudaSuccess) {
      status = fail(SODA_CUDA_CUDA_ERROR, "cudaStreamCreate failed");
      break;
    }
    status = soda_plan_prepare_host(plans[r], in_ptrs, in_strides, out_ptrs,
                                    out_strides, host_extent);
    if (status != SODA_CUDA_OK) break;
    HostPipeline& pipe = pipes[r];
    pipe.plan = plans[r];
    pipe.in_ptrs = in_ptrs;
    pipe.in_strides = in_strides;
    pipe.out_ptrs = out_ptrs;
    pipe.out_strides = out_strides;
    pipe.own_lo = bounds[r] - local_begin;
    pipe.own_hi = bounds[r + 1] - local_begin;
    pipe.host_lo = 0;                        // the ghost slices come from the
    pipe.host_hi = local_end - local_begin;  // same host arrays: no exchange
    pipe.host_shift = local_begin;
    issued[r] = pipe.issue();
  }
  for (int r = 0; r < world; ++r) {
    if (plans[r] == nullptr) continue;
    cudaSetDevice(plans[r]->device);
    int finished = pipes[r].finish(issued[r]);
    if (status == SODA_CUDA_OK) status = finished;
    cudaStream_t own = plans[r]->stream;
    plans[r]->stream = nullptr;
    soda_cuda_plan_destroy(plans[r]);
    if (own) cudaStreamDestroy(own);
  }
  cudaSetDevice(previous);
  return status;
}

int soda_cuda_host_alloc(void** ptr, int64_t bytes, int32_t device) {
  using namespace soda::rt;
  if (ptr == nullptr || bytes <= 0) return fail(SODA_CUDA_BAD_ARGUMENT, "bad size");
#ifdef SODA_EMU
  (void)device;
  *ptr = malloc(static_cast<size_t>(bytes));
  return *ptr ? SODA_CUDA_OK : fail(SODA_CUDA_OUT_OF_MEMORY, "malloc failed");
#else
  // NUMA node of the GPU: /sys/bus/pci/devices/<domain:bus:device.function>/numa_node
  int node = -1;
  if (device >= 0) {
    char bus_id[32] = {0};
    if (cudaDeviceGetPCIBusId(bus_id, sizeof(bus_id), device) == cudaSuccess) {
      for (char* c = bus_id; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = static_cast<char>(*c - 'A' + 'a');
      const std::string path = std::string("/sys/bus/pci/devices/") + bus_id + "/numa_node";
      if (FILE* fp = fopen(path.c_str(), "r")) {
        if (fscanf(fp, "%d", &node) != 1) node = -1;
        fclose(fp);
      }
    }
  }
  const size_t length = (static_cast<size_t>(bytes) + 4095) / 4096 * 4096;
  void* mem = mmap(nullptr, length, PROT_READ | PROT_WRITE,
                   MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (mem == MAP_FAILED) return fail(SODA_CUDA_OUT_OF_MEMORY, "mmap failed");
#ifdef SYS_mbind
  if (node >= 0 && node < 64) {
    // MPOL_PREFERRED: first touch lands on the GPU's node when it has room
    unsigned long mask = 1UL << node;
    syscall(SYS_mbind, mem, length, /*MPOL_PREFERRED*/ 1, &mask, 65UL, 0U);
  }
#endif
  memset(mem, 0, length);  // first touch
  cudaError_t err = cudaHostRegister(mem, length, cudaHostRegisterPortable);
  if (err != cudaSuccess) {
    munmap(mem, length);
    return fail(SODA_CUDA_OUT_OF_MEMORY,
                std::string("cudaHostRegister: ") + cudaGetErrorString(err));
  }
  *ptr = mem;
  return SODA_CUDA_OK;
#endif
}

int soda_cuda_host_free(void* ptr, int64_t bytes) {
  if (ptr == nullptr) return SODA_CUDA_OK;
#ifdef SODA_EMU
  (void)bytes;
  free(ptr);
#else
  cudaHostUnregister(ptr);
  munmap(ptr, (static_cast<size_t>(bytes) + 4095) / 4096 * 4096);
#endif
  return SODA_CUDA_OK;
}

}  // extern "C"
