// Data-parallel collectives over NCCL (NVLink 5 / NVSwitch).  The reference is single-device
// (CFFM.py:19); this is new.  libnccl is resolved at run time (dlopen) so that the single-GPU
// path and the CPU-side loader have no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <string>

#include "common.cuh"
#include "model.h"

namespace cffm {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

static NcclApi g_nccl;
static std::string g_nccl_err;

static bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
  if (!lib) { g_nccl_err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
#define SYM(field, name)                                                   \
  *(void**)(&g_nccl.field) = dlsym(lib, name);                             \
  if (!g_nccl.field) { g_nccl_err = std::string("missing symbol ") + name; dlclose(lib); return false; }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
  g_nccl.lib = lib;
  return true;
}

struct Comm { ncclComm_t comm = nullptr; };

int comm_allreduce_f32(Model* m, float* buf, int64_t n, cudaStream_t s) {
  if (!m->comm) { m->err = "communicator not initialised"; return CFFM_ERR_COMM; }
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, m->comm->comm, s);
  if (r != ncclSuccess) { m->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return CFFM_ERR_COMM; }
  m->launches++;
  return CFFM_OK;
}

int comm_allgather(Model* m, const void* send, void* recv, int64_t bytes_per_rank, cudaStream_t s) {
  if (!m->comm) { m->err = "communicator not initialised"; return CFFM_ERR_COMM; }
  ncclResult_t r = g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, ncclInt8, m->comm->comm, s);
  if (r != ncclSuccess) { m->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return CFFM_ERR_COMM; }
  m->launches++;
  return CFFM_OK;
}

// Point-to-point halves of the all-to-alls of the row-sharded tables (shard.cu); always issued inside a group.
int comm_send(Model* m, const void* buf, int64_t bytes, int peer, cudaStream_t s) {
  if (!m->comm) { m->err = "communicator not initialised"; return CFFM_ERR_COMM; }
  ncclResult_t r = g_nccl.Send(buf, (size_t)bytes, ncclInt8, peer, m->comm->comm, s);
  if (r != ncclSuccess) { m->err = std::string("ncclSend: ") + g_nccl.GetErrorString(r); return CFFM_ERR_COMM; }
  return CFFM_OK;
}
int comm_recv(Model* m, void* buf, int64_t bytes, int peer, cudaStream_t s) {
  if (!m->comm) { m->err = "communicator not initialised"; return CFFM_ERR_COMM; }
  ncclResult_t r = g_nccl.Recv(buf, (size_t)bytes, ncclInt8, peer, m->comm->comm, s);
  if (r != ncclSuccess) { m->err = std::string("ncclRecv: ") + g_nccl.GetErrorString(r); return CFFM_ERR_COMM; }
  return CFFM_OK;
}

// The collectives of one step (dense all-reduce + the row all-gathers) go out as ONE NCCL group: one fused launch
// instead of five back-to-back kernels.
int comm_group_begin(Model* m) {
  if (!m->comm) { m->err = "communicator not initialised"; return CFFM_ERR_COMM; }
  ncclResult_t r = g_nccl.GroupStart();
  if (r != ncclSuccess) { m->err = std::string("ncclGroupStart: ") + g_nccl.GetErrorString(r); return CFFM_ERR_COMM; }
  return CFFM_OK;
}
int comm_group_end(Model* m) {
  ncclResult_t r = g_nccl.GroupEnd();
  if (r != ncclSuccess) { m->err = std::string("ncclGroupEnd: ") + g_nccl.GetErrorString(r); return CFFM_ERR_COMM; }
  m->launches++;   // a group goes out as one fused kernel
  return CFFM_OK;
}

void comm_destroy(Model* m) {
  if (m->comm) {
    if (m->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm->comm);
    delete m->comm;
    m->comm = nullptr;
  }
}

}  // namespace cffm

using namespace cffm;

extern "C" int cffm_comm_unique_id(char id_out[128]) {
  if (!id_out) return CFFM_ERR_INVALID;
  if (!nccl_load()) return CFFM_ERR_COMM;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) { g_nccl_err = g_nccl.GetErrorString(r); return CFFM_ERR_COMM; }
  memcpy(id_out, &id, 128);
  return CFFM_OK;
}

extern "C" int cffm_comm_init(cffm_handle* h, const char id[128], int32_t rank, int32_t world) {
  if (!h || !id || world < 1 || rank < 0 || rank >= world) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  if (m->train_ready) { m->err = "cffm_comm_init must precede the first training step"; return CFFM_ERR_INVALID; }
  if (!nccl_load()) { m->err = g_nccl_err; return CFFM_ERR_COMM; }
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  comm_destroy(m);
  m->comm = new Comm();
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclResult_t r = g_nccl.CommInitRank(&m->comm->comm, world, uid, rank);
  if (r != ncclSuccess) {
    m->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
    delete m->comm; m->comm = nullptr;
    return CFFM_ERR_COMM;
  }
  m->world = world; m->rank = rank;
  if (sharded(m) && (world != m->shard_world || rank != m->shard_rank)) {
    m->err = "cffm_comm_init: rank / world differ from cfg.shard_rank / cfg.shard_world";
    return CFFM_ERR_INVALID;
  }
  return CFFM_OK;
}
