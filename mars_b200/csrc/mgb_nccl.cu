// libmarsgb: multi-GPU exchange of shuffle blocks -- grouped ncclSend/ncclRecv all-to-all over NVLink.
//
// Replaces the reference's storage-service shuffle transport: mapper blocks written by
// SubtaskProcessor._store_mapper_data (mars/services/subtask/worker/processor.py:396-432) and pulled by the
// reducer through StorageHandlerActor.fetch_batch / SenderManagerActor.send_batch_data in 5 MiB blocks
// (mars/services/storage/handler.py:500-619, transfer.py:63-206).  Here partial rows never leave HBM: each
// rank holds its partials partitioned by destination rank (mgb_partition_scatter) and one NCCL group moves
// every column to every peer.
//
// libnccl is dlopen()ed (the torch-bundled libnccl.so.2) so that libmarsgb.so has no link-time dependency
// on it and loads on a machine without NCCL for the single-GPU path.
#include <dlfcn.h>

#include "mgb_common.cuh"

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
#define NCCL_UINT8 1
#define NCCL_INT64 4

struct NcclApi {
  void* handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl = {nullptr};

struct mgb_comm {
  ncclComm_t comm;
  int rank, n_ranks;
};

#define MGB_NCCL(call)                                                                        \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != 0) {                                                                           \
      mgb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                             \
                    g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");       \
      return MGB_ERR_NCCL;                                                                    \
    }                                                                                         \
  } while (0)

extern "C" int mgb_nccl_load(const char* path) {
  if (g_nccl.handle) return MGB_OK;
  const char* candidates[] = {path, "libnccl.so.2", "libnccl.so", nullptr};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; ++i)
    if (candidates[i]) h = dlopen(candidates[i], RTLD_NOW | RTLD_GLOBAL);
  MGB_REQUIRE(h, MGB_ERR_NCCL, "cannot dlopen libnccl (%s): %s", path ? path : "default search", dlerror());
#define SYM(field, name)                                                            \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                        \
  MGB_REQUIRE(g_nccl.field, MGB_ERR_NCCL, "libnccl lacks symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return MGB_OK;
}

extern "C" int mgb_comm_unique_id(uint8_t out_id[128]) {
  MGB_REQUIRE(g_nccl.handle, MGB_ERR_STATE, "mgb_nccl_load was not called");
  MGB_REQUIRE(out_id, MGB_ERR_INVALID, "null pointer");
  ncclUniqueId id;
  MGB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_id, id.internal, 128);
  return MGB_OK;
}

extern "C" int mgb_comm_init(mgb_comm** out, const uint8_t id_bytes[128], int32_t rank, int32_t n_ranks) {
  MGB_REQUIRE(g_nccl.handle, MGB_ERR_STATE, "mgb_nccl_load was not called");
  MGB_REQUIRE(out && id_bytes, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, MGB_ERR_INVALID, "rank %d of %d", rank, n_ranks);
  ncclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  mgb_comm* c = new mgb_comm();
  c->rank = rank;
  c->n_ranks = n_ranks;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != 0) {
    mgb_set_error("ncclCommInitRank(rank=%d, n=%d) -> %s", rank, n_ranks, g_nccl.GetErrorString(r));
    delete c;
    return MGB_ERR_NCCL;
  }
  *out = c;
  return MGB_OK;
}

extern "C" int mgb_comm_alltoallv(mgb_comm* c, const void* const* send_cols, const int64_t* send_off,
                                  void* const* recv_cols, const int64_t* recv_off, int32_t n_cols,
                                  void* stream) {
  MGB_REQUIRE(c && send_off && recv_off, MGB_ERR_INVALID, "null pointer");
  MGB_REQUIRE(n_cols >= 0, MGB_ERR_INVALID, "negative n_cols");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = c->n_ranks;
  for (int p = 0; p < P; ++p)
    MGB_REQUIRE(send_off[p + 1] >= send_off[p] && recv_off[p + 1] >= recv_off[p], MGB_ERR_INVALID,
                "offsets must be non-decreasing");
  MGB_NCCL(g_nccl.GroupStart());
  for (int col = 0; col < n_cols; ++col) {
    const char* s = (const char*)send_cols[col];
    char* r = (char*)recv_cols[col];
    for (int p = 0; p < P; ++p) {
      const int64_t ns = send_off[p + 1] - send_off[p], nr = recv_off[p + 1] - recv_off[p];
      if (p == c->rank) {
        // own block: plain device copy, no NIC/NVLink traffic
        if (ns != nr) {
          g_nccl.GroupEnd();
          mgb_set_error("self block size mismatch: send %lld recv %lld", (long long)ns, (long long)nr);
          return MGB_ERR_INVALID;
        }
        if (ns > 0) {
          cudaError_t e = cudaMemcpyAsync(r + recv_off[p] * 8, s + send_off[p] * 8, (size_t)ns * 8,
                                          cudaMemcpyDeviceToDevice, st);
          if (e != cudaSuccess) {
            g_nccl.GroupEnd();
            mgb_set_error("self copy: %s", cudaGetErrorString(e));
            return MGB_ERR_CUDA;
          }
        }
        continue;
      }
      if (ns > 0) {
        ncclResult_t rr = g_nccl.Send(s + send_off[p] * 8, (size_t)ns, NCCL_INT64, p, c->comm, st);
        if (rr != 0) {
          g_nccl.GroupEnd();
          mgb_set_error("ncclSend -> %s", g_nccl.GetErrorString(rr));
          return MGB_ERR_NCCL;
        }
      }
      if (nr > 0) {
        ncclResult_t rr = g_nccl.Recv(r + recv_off[p] * 8, (size_t)nr, NCCL_INT64, p, c->comm, st);
        if (rr != 0) {
          g_nccl.GroupEnd();
          mgb_set_error("ncclRecv -> %s", g_nccl.GetErrorString(rr));
          return MGB_ERR_NCCL;
        }
      }
    }
  }
  MGB_NCCL(g_nccl.GroupEnd());
  return MGB_OK;
}

extern "C" int mgb_comm_allgather_i64(mgb_comm* c, const int64_t* send, int64_t* recv, int64_t count,
                                      void* stream) {
  MGB_REQUIRE(c && send && recv && count >= 0, MGB_ERR_INVALID, "bad argument");
  MGB_NCCL(g_nccl.AllGather(send, recv, (size_t)count, NCCL_INT64, c->comm, (cudaStream_t)stream));
  return MGB_OK;
}

extern "C" int mgb_comm_destroy(mgb_comm* c) {
  if (!c) return MGB_OK;
  if (g_nccl.handle && c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return MGB_OK;
}


// ----------------------------------------------------------------------------------------------------
// Peer-to-peer pull transport (alternative to the NCCL all-to-all on one NVSwitch box): every rank packs its
// send buffer into a library-owned arena, publishes the arena's CUDA IPC handle together with the block
// counts, and each receiver pulls its slices out of the peers' arenas with cudaMemcpyAsync -- the copy
// engines move the bytes over NVLink, no SM and no NCCL kernel involved.
// ----------------------------------------------------------------------------------------------------
struct IpcArena {
  void* base = nullptr;
  size_t bytes = 0;
  cudaIpcMemHandle_t handle;
};
struct IpcPeer {
  bool open = false;
  cudaIpcMemHandle_t handle;
  void* ptr = nullptr;
};
#define MGB_IPC_SLOTS 4
static IpcArena g_arena[8][MGB_IPC_SLOTS];          // [local device][slot]
static IpcPeer g_peers[8][64][MGB_IPC_SLOTS];       // [local device][peer rank][slot]

static int ipc_device(int* dev) {
  MGB_CUDA(cudaGetDevice(dev));
  MGB_REQUIRE(*dev >= 0 && *dev < 8, MGB_ERR_UNSUPPORTED, "device ordinal %d", *dev);
  return MGB_OK;
}

extern "C" int mgb_ipc_arena_slot(int32_t slot, int64_t bytes, void** out_ptr, uint8_t out_handle[64],
                                  int64_t* out_capacity, int32_t* out_fresh) {
  MGB_REQUIRE(out_ptr && out_handle && bytes >= 0 && slot >= 0 && slot < MGB_IPC_SLOTS, MGB_ERR_INVALID, "bad argument");
  int dev = 0;
  MGB_TRY(ipc_device(&dev));
  IpcArena& a = g_arena[dev][slot];
  if (out_fresh) *out_fresh = 0;
  if (a.bytes < (size_t)bytes || !a.base) {
    // the caller guarantees that no peer still maps the old allocation (close on the peers + barrier, see mgb.h)
    if (a.base) MGB_CUDA(cudaFree(a.base));
    size_t want = (size_t)bytes + ((size_t)bytes >> 2);
    if (want < (64u << 20)) want = 64u << 20;
    a.base = nullptr;
    a.bytes = 0;
    MGB_CUDA(cudaMalloc(&a.base, want));
    a.bytes = want;
    MGB_CUDA(cudaIpcGetMemHandle(&a.handle, a.base));
    if (out_fresh) *out_fresh = 1;
  }
  *out_ptr = a.base;
  if (out_capacity) *out_capacity = (int64_t)a.bytes;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(out_handle, &a.handle, 64);
  return MGB_OK;
}

extern "C" int mgb_ipc_arena(int64_t bytes, void** out_ptr, uint8_t out_handle[64]) {
  return mgb_ipc_arena_slot(0, bytes, out_ptr, out_handle, nullptr, nullptr);
}

extern "C" int mgb_ipc_close_slot(int32_t peer, int32_t slot) {
  MGB_REQUIRE(peer >= 0 && peer < 64 && slot >= 0 && slot < MGB_IPC_SLOTS, MGB_ERR_INVALID, "bad argument");
  int dev = 0;
  MGB_TRY(ipc_device(&dev));
  IpcPeer& p = g_peers[dev][peer][slot];
  if (p.open) {
    MGB_CUDA(cudaIpcCloseMemHandle(p.ptr));
    p.open = false;
    p.ptr = nullptr;
  }
  return MGB_OK;
}

extern "C" int mgb_ipc_open_slot(int32_t peer, int32_t slot, const uint8_t handle[64], void** out_ptr) {
  MGB_REQUIRE(handle && out_ptr && peer >= 0 && peer < 64 && slot >= 0 && slot < MGB_IPC_SLOTS, MGB_ERR_INVALID,
              "bad argument");
  int dev = 0;
  MGB_TRY(ipc_device(&dev));
  IpcPeer& p = g_peers[dev][peer][slot];
  if (p.open && memcmp(&p.handle, handle, 64) == 0) {
    *out_ptr = p.ptr;
    return MGB_OK;
  }
  if (p.open) {   // the peer re-allocated its arena: drop the stale mapping
    cudaIpcCloseMemHandle(p.ptr);
    p.open = false;
  }
  memcpy(&p.handle, handle, 64);
  MGB_CUDA(cudaIpcOpenMemHandle(&p.ptr, p.handle, cudaIpcMemLazyEnablePeerAccess));
  p.open = true;
  *out_ptr = p.ptr;
  return MGB_OK;
}

extern "C" int mgb_ipc_open(int32_t peer, const uint8_t handle[64], void** out_ptr) {
  return mgb_ipc_open_slot(peer, 0, handle, out_ptr);
}

extern "C" int mgb_ipc_pull(void* dst, const void* peer_src, int64_t bytes, void* stream) {
  MGB_REQUIRE(bytes >= 0, MGB_ERR_INVALID, "negative size");
  if (bytes == 0) return MGB_OK;
  MGB_REQUIRE(dst && peer_src, MGB_ERR_INVALID, "null pointer");
  MGB_CUDA(cudaMemcpyAsync(dst, peer_src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return MGB_OK;
}
