// Communicators of the sharded step (include/fvx.h: FvxComm): two NCCL communicators over the same ranks,
// one used on the caller's stream and one on a side stream.  NCCL is resolved at run time with dlopen /
// dlsym - the copy the process already holds (torch's bundled libnccl.so.2) if there is one, else the system
// library - so libfvx.so has no link-time dependency on it and a single-GPU user never loads it.
#include <dlfcn.h>
#include <nccl.h>      // types and enums only (compile time); every call goes through the table below
#include <string.h>

#include "fvx_common.cuh"
#include "fvx_comm.cuh"

struct NcclApi {
  void* handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // already in the process (torch)?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) FVX_FAIL(-5, "fvx_comm: cannot load libnccl.so.2: %s", dlerror());
#define FVX_SYM(field, name)                                                       \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                       \
  if (!g_nccl.field) FVX_FAIL(-5, "fvx_comm: libnccl has no symbol %s", name);
  FVX_SYM(GetUniqueId, "ncclGetUniqueId")
  FVX_SYM(CommInitRank, "ncclCommInitRank")
  FVX_SYM(CommDestroy, "ncclCommDestroy")
  FVX_SYM(AllReduce, "ncclAllReduce")
  FVX_SYM(AllGather, "ncclAllGather")
  FVX_SYM(ReduceScatter, "ncclReduceScatter")
  FVX_SYM(GroupStart, "ncclGroupStart")
  FVX_SYM(GroupEnd, "ncclGroupEnd")
  FVX_SYM(GetErrorString, "ncclGetErrorString")
#undef FVX_SYM
  g_nccl.handle = h;
  return 0;
}

#define FVX_NCCL(call, what)                                                                  \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != ncclSuccess) FVX_FAIL(-5, "%s: NCCL error: %s", (what), g_nccl.GetErrorString(r__)); \
  } while (0)

int fvx_comm_allreduce(FvxComm* c, int which, float* buf, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  FVX_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat, ncclSum, reinterpret_cast<ncclComm_t>(c->nccl[which]), st),
           "all-reduce");
  return 0;
}

// in place over R equal segments of `seg` floats at `buf`: rank r's segment is buf + r * seg
int fvx_comm_allgather(FvxComm* c, int which, float* buf, size_t seg, cudaStream_t st) {
  if (seg == 0) return 0;
  FVX_NCCL(g_nccl.AllGather(buf + (size_t)c->rank * seg, buf, seg, ncclFloat, reinterpret_cast<ncclComm_t>(c->nccl[which]), st),
           "all-gather");
  return 0;
}
int fvx_comm_reducescatter(FvxComm* c, int which, float* buf, size_t seg, cudaStream_t st) {
  if (seg == 0) return 0;
  FVX_NCCL(g_nccl.ReduceScatter(buf, buf + (size_t)c->rank * seg, seg, ncclFloat, ncclSum,
                                reinterpret_cast<ncclComm_t>(c->nccl[which]), st), "reduce-scatter");
  return 0;
}

extern "C" {

int fvx_comm_unique_id(uint8_t* id_host) {
  FVX_CHECK_ARG(id_host != nullptr, "fvx_comm_unique_id: null buffer");
  if (int rc = nccl_load()) return rc;
  static_assert(2 * sizeof(ncclUniqueId) <= FVX_COMM_ID_BYTES, "id buffer too small");
  memset(id_host, 0, FVX_COMM_ID_BYTES);
  for (int i = 0; i < 2; ++i) {
    ncclUniqueId id;
    FVX_NCCL(g_nccl.GetUniqueId(&id), "fvx_comm_unique_id");
    memcpy(id_host + i * sizeof(ncclUniqueId), &id, sizeof(id));
  }
  return 0;
}

int fvx_comm_create(const uint8_t* id_host, int32_t rank, int32_t world, FvxComm** out) {
  FVX_CHECK_ARG(id_host && out && world >= 1 && rank >= 0 && rank < world, "fvx_comm_create: bad arguments");
  if (int rc = nccl_load()) return rc;
  FvxComm* c = new FvxComm();
  memset(c, 0, sizeof(*c));
  c->rank = rank;
  c->world = world;
  for (int i = 0; i < 2; ++i) {
    ncclUniqueId id;
    memcpy(&id, id_host + i * sizeof(ncclUniqueId), sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) {
      fvx_set_error("fvx_comm_create: ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
      delete c;
      return -5;
    }
    c->nccl[i] = comm;
  }
  bool ok = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < FVX_COMM_EVENTS && ok; ++i) ok = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    fvx_set_error("fvx_comm_create: cannot create the side stream / events: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return -3;
  }
  *out = c;
  return 0;
}

// Peer-mapped arena: every rank allocates `bytes`, the CUDA IPC handles travel through an all-gather on the
// communicator, and each rank maps its peers' allocations (NVLink peer access is enabled by the mapping).
int fvx_comm_arena(FvxComm* c, int64_t bytes, void** local_out) {
  FVX_CHECK_ARG(c && local_out && bytes > 0, "fvx_comm_arena: bad arguments");
  FVX_CHECK_ARG(c->world <= FVX_COMM_MAX_RANKS, "fvx_comm_arena: at most %d ranks", FVX_COMM_MAX_RANKS);
  if (c->arena) {
    // a new arena replaces the old one (whoever used it must be done): every rank drains its device, the ranks
    // meet in a collective, then the mappings and the allocation go
    cudaDeviceSynchronize();
    float* tmp = nullptr;
    if (cudaMalloc(&tmp, 256) != cudaSuccess) FVX_FAIL(-3, "fvx_comm_arena: cudaMalloc failed");
    cudaMemset(tmp, 0, 256);
    FVX_NCCL(g_nccl.AllReduce(tmp, tmp, 1, ncclFloat, ncclSum, reinterpret_cast<ncclComm_t>(c->nccl[0]), 0), "fvx_comm_arena");
    cudaDeviceSynchronize();
    cudaFree(tmp);
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->arena);
    c->arena = nullptr;
  }
  // a rank whose allocation or export fails still takes part in the exchange (with ok = 0), so that every rank
  // leaves this call with the same verdict instead of waiting in the collective for the one that gave up
  struct Slot { cudaIpcMemHandle_t h; int32_t ok, pad[3]; };
  static_assert(sizeof(Slot) % 4 == 0, "slot size");
  Slot mine;
  memset(&mine, 0, sizeof(mine));
  void* base = nullptr;
  const char* why = nullptr;
  if (cudaMalloc(&base, (size_t)bytes) != cudaSuccess) { why = "cudaMalloc"; base = nullptr; }
  else if (cudaMemset(base, 0, (size_t)bytes) != cudaSuccess) why = "cudaMemset";
  else if (cudaIpcGetMemHandle(&mine.h, base) != cudaSuccess) why = "cudaIpcGetMemHandle";
  char why_msg[256] = "";
  if (why) snprintf(why_msg, sizeof(why_msg), "%s(%lld bytes): %s", why, (long long)bytes, cudaGetErrorString(cudaGetLastError()));
  mine.ok = why ? 0 : 1;
  const size_t hs = sizeof(Slot);
  uint8_t* dev = nullptr;
  if (cudaMalloc(&dev, hs * c->world) != cudaSuccess) FVX_FAIL(-3, "fvx_comm_arena: cudaMalloc failed");
  cudaMemcpy(dev + hs * c->rank, &mine, hs, cudaMemcpyHostToDevice);
  FVX_NCCL(g_nccl.AllGather(dev + hs * c->rank, dev, hs / 4, ncclFloat, reinterpret_cast<ncclComm_t>(c->nccl[0]), 0),
           "fvx_comm_arena");
  if (cudaDeviceSynchronize() != cudaSuccess)
    FVX_FAIL(-3, "fvx_comm_arena: handle exchange failed: %s", cudaGetErrorString(cudaGetLastError()));
  Slot all[FVX_COMM_MAX_RANKS];
  cudaMemcpy(all, dev, hs * c->world, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  for (int r = 0; r < c->world; ++r)
    if (!all[r].ok) {
      if (base) cudaFree(base);
      if (r == c->rank) FVX_FAIL(-3, "fvx_comm_arena: %s", why_msg);
      FVX_FAIL(-3, "fvx_comm_arena: rank %d could not allocate or export its arena", r);
    }
  for (int r = 0; r < c->world; ++r) c->peer[r] = nullptr;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer[r] = reinterpret_cast<uint8_t*>(base); continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      const cudaError_t e = cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (q != c->rank && c->peer[q]) { cudaIpcCloseMemHandle(c->peer[q]); c->peer[q] = nullptr; }
      c->peer[c->rank] = nullptr;
      // (the allocation stays alive until the communicator goes: a peer may have mapped it already)
      c->arena = reinterpret_cast<uint8_t*>(base);
      c->arena_bytes = (size_t)bytes;
      FVX_FAIL(-3, "fvx_comm_arena: cannot map the arena of rank %d: %s", r, cudaGetErrorString(e));
    }
    c->peer[r] = reinterpret_cast<uint8_t*>(p);
  }
  c->arena = reinterpret_cast<uint8_t*>(base);
  c->arena_bytes = (size_t)bytes;
  c->epoch = 0;
  *local_out = base;
  return 0;
}

int fvx_comm_destroy(FvxComm* c) {
  if (!c) return 0;
  if (c->arena) {
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->arena);
  }
  for (int i = 0; i < 2; ++i)
    if (c->nccl[i]) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(c->nccl[i]));
  if (c->side) cudaStreamDestroy(c->side);
  for (int i = 0; i < FVX_COMM_EVENTS; ++i)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  delete c;
  return 0;
}

int fvx_comm_all_reduce_f32(FvxComm* c, float* buf, int64_t n, fvx_stream_t stream) {
  FVX_CHECK_ARG(c && buf && n >= 0, "fvx_comm_all_reduce_f32: bad arguments");
  return fvx_comm_allreduce(c, 0, buf, (size_t)n, fvx_cu(stream));
}

}  // extern "C"
