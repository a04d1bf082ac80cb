// NCCL collectives over NVLink 5 / NVSwitch behind the C-ABI (one communicator per process/GPU).
// Replaces distributed/communication.py:115-347 of the reference (CuPy's NCCL binding, whose unique-id
// exchange is a stub: non-zero ranks call get_unique_id() themselves, :339-347, so it can never rendezvous).
// Here rank 0 calls nfb_comm_unique_id and the HOST code ships the 128 bytes to the other ranks (TCP store /
// torch.distributed gloo / a file); every rank then calls nfb_comm_init.  libnccl.so.2 is dlopen'ed at run time
// (the pip wheel nvidia-nccl-cu12 2.28.9 in this image), so libnfb200.so has no link-time dependency on it.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

static void* g_nccl = nullptr;
static ncclResult_t (*p_GetUniqueId)(ncclUniqueId*) = nullptr;
static ncclResult_t (*p_CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
static ncclResult_t (*p_CommDestroy)(ncclComm_t) = nullptr;
static ncclResult_t (*p_AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
static ncclResult_t (*p_Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
static ncclResult_t (*p_AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
static ncclResult_t (*p_ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
static const char* (*p_GetErrorString)(ncclResult_t) = nullptr;

#define NFB_NCCL(expr)                                                                                       \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) return NFB_FAIL("%s -> %s", #expr, p_GetErrorString ? p_GetErrorString(_r) : "nccl error"); \
  } while (0)

NFB_API int nfb_comm_load(const char* path) {
  if (g_nccl) return 0;
  const char* cands[3] = {path && path[0] ? path : nullptr, "libnccl.so.2", "libnccl.so"};
  for (int i = 0; i < 3 && !g_nccl; ++i)
    if (cands[i]) g_nccl = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl) return NFB_FAIL("cannot dlopen libnccl.so.2 (%s)", dlerror());
#define SYM(var, name)                                        \
  *(void**)(&var) = dlsym(g_nccl, name);                      \
  if (!var) return NFB_FAIL("libnccl is missing symbol %s", name);
  SYM(p_GetUniqueId, "ncclGetUniqueId")
  SYM(p_CommInitRank, "ncclCommInitRank")
  SYM(p_CommDestroy, "ncclCommDestroy")
  SYM(p_AllReduce, "ncclAllReduce")
  SYM(p_Broadcast, "ncclBroadcast")
  SYM(p_AllGather, "ncclAllGather")
  SYM(p_ReduceScatter, "ncclReduceScatter")
  SYM(p_GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}

NFB_API int nfb_comm_unique_id(char* out, int len) {
  if (!g_nccl) return NFB_FAIL("nfb_comm_load was not called");
  if (len < (int)sizeof(ncclUniqueId)) return NFB_FAIL("unique id buffer too small (%d < %zu)", len, sizeof(ncclUniqueId));
  ncclUniqueId id;
  NFB_NCCL(p_GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return 0;
}

NFB_API int nfb_comm_init(int rank, int world, const char* id_bytes, int len, void** comm_out) {
  if (!g_nccl) return NFB_FAIL("nfb_comm_load was not called");
  if (len < (int)sizeof(ncclUniqueId)) return NFB_FAIL("unique id too short");
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm;
  NFB_NCCL(p_CommInitRank(&comm, world, id, rank));
  *comm_out = (void*)comm;
  return 0;
}

NFB_API int nfb_comm_destroy(void* comm) {
  if (comm && p_CommDestroy) NFB_NCCL(p_CommDestroy((ncclComm_t)comm));
  return 0;
}

static int nccl_dtype(int dt, ncclDataType_t* out) {
  switch (dt) {
    case NFB_F32: *out = ncclFloat32; return 0;
    case NFB_F64: *out = ncclFloat64; return 0;
    case NFB_I32: *out = ncclInt32; return 0;
    case NFB_I64: *out = ncclInt64; return 0;
    case NFB_BOOL: *out = ncclUint8; return 0;
    case NFB_BF16: *out = ncclBfloat16; return 0;
  }
  return NFB_FAIL("unsupported dtype %d for a collective", dt);
}
// op: 0 sum, 1 average, 2 max, 3 min, 4 product   (distributed/communication.py:25-32 ReduceOp)
static int nccl_op(int op, ncclRedOp_t* out) {
  switch (op) {
    case 0: *out = ncclSum; return 0;
    case 1: *out = ncclAvg; return 0;
    case 2: *out = ncclMax; return 0;
    case 3: *out = ncclMin; return 0;
    case 4: *out = ncclProd; return 0;
  }
  return NFB_FAIL("unsupported reduce op %d", op);
}

NFB_API int nfb_comm_allreduce(void* comm, const void* send, void* recv, int64_t count, int dtype, int op, void* stream) {
  ncclDataType_t dt;
  ncclRedOp_t ro;
  if (nccl_dtype(dtype, &dt) || nccl_op(op, &ro)) return -1;
  NFB_NCCL(p_AllReduce(send, recv, (size_t)count, dt, ro, (ncclComm_t)comm, stream ? (cudaStream_t)stream : nfb_cur_stream()));
  return 0;
}
NFB_API int nfb_comm_broadcast(void* comm, void* buf, int64_t count, int dtype, int root, void* stream) {
  ncclDataType_t dt;
  if (nccl_dtype(dtype, &dt)) return -1;
  NFB_NCCL(p_Broadcast(buf, buf, (size_t)count, dt, root, (ncclComm_t)comm, stream ? (cudaStream_t)stream : nfb_cur_stream()));
  return 0;
}
NFB_API int nfb_comm_allgather(void* comm, const void* send, void* recv, int64_t sendcount, int dtype, void* stream) {
  ncclDataType_t dt;
  if (nccl_dtype(dtype, &dt)) return -1;
  NFB_NCCL(p_AllGather(send, recv, (size_t)sendcount, dt, (ncclComm_t)comm, stream ? (cudaStream_t)stream : nfb_cur_stream()));
  return 0;
}
NFB_API int nfb_comm_reduce_scatter(void* comm, const void* send, void* recv, int64_t recvcount, int dtype, int op, void* stream) {
  ncclDataType_t dt;
  ncclRedOp_t ro;
  if (nccl_dtype(dtype, &dt) || nccl_op(op, &ro)) return -1;
  NFB_NCCL(p_ReduceScatter(send, recv, (size_t)recvcount, dt, ro, (ncclComm_t)comm, stream ? (cudaStream_t)stream : nfb_cur_stream()));
  return 0;
}
