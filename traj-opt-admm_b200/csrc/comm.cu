// comm.cu -- multi-GPU exchange of the multi-robot path, native NCCL on the context's stream.
//
// Robots are block-partitioned over the ranks of one communicator (one process or host thread per GPU); the cloud and
// its LBVH are replicated.  What crosses GPUs per ADMM iteration are robot-indexed arrays whose
// owned slice is valid on each rank:
//   decoupled (Optimization3D_multi::optimization_decouple, Optimization3D_multi.h:29-118)
//     control points before separate_self (:51), directions + wolfe + gnorm before Step::self_step (:72-76)
//   coupled (::optimization -> update_spline :508-639)
//     the seven Schur sums of every robot's block (:519-557), directions, the CCD ladder exponents (min step, :586-594)
//     and the trial energies of every Armijo round (:605-636)
// Every exchange is an in-place all-gather (equal shares) or one grouped broadcast per rank (unequal shares); sums over
// robots are then taken locally in robot order, so a sharded run is bitwise equal to a single context.  The calls are
// issued on c->stream and are captured into the iteration's CUDA graph like any kernel.
//
// NCCL is bound at run time (dlopen of libnccl.so.2; an already loaded copy -- e.g. the one bundled with torch -- is
// reused) so that the library has no load-time dependency on it: single-GPU users never touch NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include <mutex>

#include "ctx.cuh"

namespace tob {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* names[4] = {getenv("TRAJOPT_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
    for (int pass = 0; pass < 2 && !api.handle; pass++)      // pass 0: a copy that is already loaded (torch's), pass 1: load
      for (int i = 0; i < 3 && !api.handle; i++)
        if (names[i] && names[i][0]) api.handle = dlopen(names[i], RTLD_NOW | RTLD_LOCAL | (pass == 0 ? RTLD_NOLOAD : 0));
    if (!api.handle) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
#define TOB_SYM(field, name)                                                       \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));      \
  if (!api.field) { api.err = std::string("NCCL symbol missing: ") + name; return; }
    TOB_SYM(GetVersion, "ncclGetVersion") TOB_SYM(GetUniqueId, "ncclGetUniqueId") TOB_SYM(CommInitRank, "ncclCommInitRank")
    TOB_SYM(CommInitAll, "ncclCommInitAll") TOB_SYM(CommDestroy, "ncclCommDestroy") TOB_SYM(CommCount, "ncclCommCount")
    TOB_SYM(CommUserRank, "ncclCommUserRank") TOB_SYM(AllGather, "ncclAllGather") TOB_SYM(Broadcast, "ncclBroadcast")
    TOB_SYM(AllReduce, "ncclAllReduce") TOB_SYM(GroupStart, "ncclGroupStart") TOB_SYM(GroupEnd, "ncclGroupEnd")
    TOB_SYM(GetErrorString, "ncclGetErrorString")
#undef TOB_SYM
  });
  return &api;
}

static int nccl_fail(tob_ctx* c, const char* what, ncclResult_t r) {
  NcclApi* n = nccl_api();
  return fail_msg(c, std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(r) : "NCCL error"));
}
#define TOB_NCCL(c, call)                                       \
  do {                                                          \
    ncclResult_t r__ = (call);                                  \
    if (r__ != ncclSuccess) return nccl_fail((c), #call, r__);  \
  } while (0)

// block partition of the robots over the ranks: the first (U mod world) ranks own one robot more
void shard_partition(tob_ctx* c) {
  const int U = c->n_robots(), W = c->comm_world, base = U / W, rem = U % W;
  c->shard_first.resize(W); c->shard_count.resize(W);
  for (int r = 0; r < W; r++) {
    c->shard_count[r] = base + (r < rem ? 1 : 0);
    c->shard_first[r] = r * base + (r < rem ? r : rem);
  }
  c->own_begin = c->shard_first[c->comm_rank];
  c->own_end = c->own_begin + c->shard_count[c->comm_rank];
}

int comm_attach(tob_ctx* c, void* comm, bool owned) {
  NcclApi* n = nccl_api();
  if (!n->handle || !n->err.empty()) return fail_msg(c, "NCCL unavailable: " + n->err);
  int world = 0, rank = 0;
  TOB_NCCL(c, n->CommCount((ncclComm_t)comm, &world));
  TOB_NCCL(c, n->CommUserRank((ncclComm_t)comm, &rank));
  if (c->have_params && world > c->n_robots()) return fail_msg(c, "more ranks than robots: every rank must own at least one robot");
  c->nccl_comm = comm; c->nccl_owned = owned; c->comm_rank = rank; c->comm_world = world;
  c->ag = nullptr; c->ar = nullptr; c->cb_user = nullptr;
  if (c->have_params) shard_partition(c);
  TOB_CUDA(c, c->ovf_all.ensure((size_t)world + 1));
  if (!c->comm_stream) {
    TOB_CUDA(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    TOB_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    TOB_CUDA(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_mode = -1; }
  return 0;
}

// fork: what follows on the exchange stream waits for everything enqueued on the main stream so far
int exchange_fork(tob_ctx* c) {
  if (!c->nccl_comm || !c->comm_stream) return 0;
  TOB_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
  TOB_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_fork, 0));
  c->xs = c->comm_stream;
  return 0;
}
// join: the main stream waits for the exchanges issued since the fork
int exchange_join(tob_ctx* c) {
  if (!c->xs) return 0;
  TOB_CUDA(c, cudaEventRecord(c->ev_join, c->comm_stream));
  TOB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  c->xs = nullptr;
  return 0;
}

void comm_release(tob_ctx* c) {
  if (c->nccl_comm && c->nccl_owned) {
    NcclApi* n = nccl_api();
    if (n->CommDestroy) n->CommDestroy((ncclComm_t)c->nccl_comm);
  }
  c->nccl_comm = nullptr; c->nccl_owned = false; c->comm_rank = 0; c->comm_world = 1;
}

int exchange_group_begin(tob_ctx* c) {
  if (c->nccl_comm) TOB_NCCL(c, nccl_api()->GroupStart());
  return 0;
}
int exchange_group_end(tob_ctx* c) {
  if (c->nccl_comm) TOB_NCCL(c, nccl_api()->GroupEnd());
  return 0;
}

// In-place exchange of a robot-indexed device array (elems_per_robot elements of elem_size bytes per robot): on return
// (stream order) every rank holds every robot's slice.  No-op for an unsharded context.
int exchange_robots(tob_ctx* c, void* buf, size_t elems_per_robot, size_t elem_size) {
  if (c->nccl_comm) {
    NcclApi* n = nccl_api();
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const ncclDataType_t dt = elem_size == 8 ? ncclFloat64 : (elem_size == 4 ? ncclInt32 : ncclInt8);
    const size_t unit = elem_size == 8 || elem_size == 4 ? 1 : elem_size;   // other sizes travel as bytes
    const size_t stride = elems_per_robot * elem_size;                      // bytes per robot
    const int W = c->comm_world;
    cudaStream_t xs = c->xs ? c->xs : c->stream;
    bool equal = true;
    for (int r = 1; r < W; r++) equal = equal && c->shard_count[r] == c->shard_count[0];
    char* base = (char*)buf;
    if (equal) {
      const size_t cnt = (size_t)c->shard_count[0] * elems_per_robot * unit;
      TOB_NCCL(c, n->AllGather(base + (size_t)c->own_begin * stride, base, cnt, dt, comm, xs));
    } else {
      TOB_NCCL(c, n->GroupStart());
      for (int r = 0; r < W; r++) {
        char* p = base + (size_t)c->shard_first[r] * stride;
        TOB_NCCL(c, n->Broadcast(p, p, (size_t)c->shard_count[r] * elems_per_robot * unit, dt, r, comm, xs));
      }
      TOB_NCCL(c, n->GroupEnd());
    }
    return 0;
  }
  if (c->ag) {
    if (elem_size != 8) return fail_msg(c, "exchange through the tob_set_shard callbacks carries FP64 only: attach NCCL (tob_nccl_*) for this mode");
    const size_t per_rank = elems_per_robot * (size_t)(c->own_end - c->own_begin);
    if (c->ag(buf, per_rank, c->cb_user)) return fail_msg(c, "all-gather callback failed");
  }
  return 0;
}

// one FP64 word per rank (rank r's word at buf[r])
int exchange_ranks(tob_ctx* c, double* buf) {
  if (c->nccl_comm) {
    NcclApi* n = nccl_api();
    TOB_NCCL(c, n->AllGather(buf + c->comm_rank, buf, 1, ncclFloat64, (ncclComm_t)c->nccl_comm, c->xs ? c->xs : c->stream));
    return 0;
  }
  if (c->ag && c->ag(buf, 1, c->cb_user)) return fail_msg(c, "all-gather callback failed");
  return 0;
}

}  // namespace tob

using namespace tob;

extern "C" {

int tob_nccl_available(int* version) {
  NcclApi* n = nccl_api();
  if (!n->handle || !n->err.empty()) return 1;
  if (version) n->GetVersion(version);
  return 0;
}

int tob_nccl_unique_id(void* id128) {
  NcclApi* n = nccl_api();
  if (!n->handle || !n->err.empty() || !id128) return fail_msg(nullptr, "NCCL unavailable: " + n->err);
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != ncclSuccess) return fail_msg(nullptr, "ncclGetUniqueId failed");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return 0;
}

int tob_nccl_init_rank(tob_ctx* c, const void* id128, int rank, int world) {
  if (!c || !id128) return 1;
  NcclApi* n = nccl_api();
  if (!n->handle || !n->err.empty()) return fail_msg(c, "NCCL unavailable: " + n->err);
  cudaSetDevice(c->device);
  comm_release(c);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  TOB_NCCL(c, n->CommInitRank(&comm, world, id, rank));
  return comm_attach(c, comm, true);
}

int tob_nccl_attach(tob_ctx* c, void* nccl_comm) {
  if (!c || !nccl_comm) return 1;
  cudaSetDevice(c->device);
  comm_release(c);
  return comm_attach(c, nccl_comm, false);
}

int tob_nccl_init_all(tob_ctx** ctxs, int n_ctx) {
  if (!ctxs || n_ctx < 1) return 1;
  NcclApi* n = nccl_api();
  if (!n->handle || !n->err.empty()) return fail_msg(ctxs[0], "NCCL unavailable: " + n->err);
  std::vector<int> devs(n_ctx);
  std::vector<ncclComm_t> comms(n_ctx);
  for (int i = 0; i < n_ctx; i++) { if (!ctxs[i]) return 1; devs[i] = ctxs[i]->device; comm_release(ctxs[i]); }
  TOB_NCCL(ctxs[0], n->CommInitAll(comms.data(), n_ctx, devs.data()));
  for (int i = 0; i < n_ctx; i++) {
    cudaSetDevice(ctxs[i]->device);
    TOB_TRY(comm_attach(ctxs[i], comms[i], true));
  }
  return 0;
}

int tob_nccl_detach(tob_ctx* c) {
  if (!c) return 1;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  comm_release(c);
  c->own_begin = 0; c->own_end = c->have_params ? c->n_robots() : 0;
  if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_mode = -1; }
  return 0;
}

int tob_shard_range(const tob_ctx* c, int* first, int* count, int* rank, int* world) {
  if (!c) return 1;
  if (first) *first = c->own_begin;
  if (count) *count = c->own_end - c->own_begin;
  if (rank) *rank = c->comm_rank;
  if (world) *world = c->comm_world;
  return 0;
}

}  // extern "C"
