// lfm_comm_*: the collective of the sharded batched path behind the C-ABI (north_star: "one NCCL
// allreduce of best-objective ... state per step over NVLink").  A thin holder of an ncclComm_t with exactly the two
// collectives the path uses: an in-place integer MIN all-reduce of the best-objective keys (best_key / step_keys of
// lfm_batched_fit_*) and an all-gather of the per-rank winners (lfm_batched_best).  NCCL is bound at RUN time
// (dlopen of libnccl.so.2 -- inside a PyTorch process that is the NCCL torch already loaded), so liblfm_b200.so keeps
// linking against libcudart only and still loads on hosts without NCCL; the entry points then return LFM_ERR_COMM.
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: every call goes through dlsym

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "lfm_common.cuh"

namespace {
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  bool ok = false;
};
Nccl g_nccl;
std::once_flag g_nccl_once;

const Nccl& nccl() {
  std::call_once(g_nccl_once, [] {
    const char* override_path = getenv("LFM_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    auto sym = [&](const char* s) { return dlsym(g_nccl.handle, s); };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.AllGather;
  });
  return g_nccl;
}
}  // namespace

struct lfm_comm {
  ncclComm_t comm;
  int world, rank, device;
};

static_assert(sizeof(ncclUniqueId) == LFM_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");

extern "C" int lfm_comm_available(void) { return nccl().ok ? 1 : 0; }

extern "C" int lfm_comm_unique_id(void* id_out) {
  if (!id_out) return LFM_ERR_INVALID;
  if (!nccl().ok) return LFM_ERR_COMM;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return LFM_ERR_COMM;
  memcpy(id_out, &id, sizeof(id));
  return LFM_OK;
}

extern "C" int lfm_comm_create(lfm_comm** out, int world, int rank, const void* id_bytes) {
  if (!out || !id_bytes || world <= 0 || rank < 0 || rank >= world) return LFM_ERR_INVALID;
  if (!nccl().ok) return LFM_ERR_COMM;
  LFM_TRY(lfm_device_check());
  lfm_comm* c = (lfm_comm*)calloc(1, sizeof(lfm_comm));
  if (!c) return LFM_ERR_INVALID;
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  c->world = world; c->rank = rank;
  cudaGetDevice(&c->device);
  if (nccl().CommInitRank(&c->comm, world, id, rank) != ncclSuccess) { free(c); return LFM_ERR_COMM; }
  *out = c;
  return LFM_OK;
}

extern "C" int lfm_comm_world(const lfm_comm* c) { return c ? c->world : 0; }
extern "C" int lfm_comm_rank(const lfm_comm* c) { return c ? c->rank : -1; }

extern "C" int lfm_comm_allreduce_min_i64(lfm_comm* c, long long* buf, size_t count, lfm_stream_t stream) {
  if (!c || (!buf && count)) return LFM_ERR_INVALID;
  if (count == 0) return LFM_OK;
  return nccl().AllReduce(buf, buf, count, ncclInt64, ncclMin, c->comm, (cudaStream_t)stream) == ncclSuccess ? LFM_OK
                                                                                                             : LFM_ERR_COMM;
}

extern "C" int lfm_comm_allgather_f64(lfm_comm* c, const double* send, double* recv, size_t count, lfm_stream_t stream) {
  if (!c || ((!send || !recv) && count)) return LFM_ERR_INVALID;
  if (count == 0) return LFM_OK;
  return nccl().AllGather(send, recv, count, ncclFloat64, c->comm, (cudaStream_t)stream) == ncclSuccess ? LFM_OK
                                                                                                        : LFM_ERR_COMM;
}

extern "C" int lfm_comm_destroy(lfm_comm* c) {
  if (!c) return LFM_OK;
  const int st = nccl().ok && nccl().CommDestroy(c->comm) == ncclSuccess ? LFM_OK : LFM_ERR_COMM;
  free(c);
  return st;
}
