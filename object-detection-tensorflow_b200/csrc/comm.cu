// Multi-GPU plumbing of the detection path behind the C ABI: one NCCL communicator per
// process (one process per GPU), ONE all-gather of the packed fixed-size detection records per batch -- the records
// are written by the NMS kernel itself (tail.cu), so the collective follows on the same stream with no pack step --
// and a broadcast for replicating weights.  NCCL is resolved at run time with dlopen (the copy PyTorch has already
// loaded is found by its soname), so the library has no link-time dependency on it and single-GPU users never
// touch it.  The reference is single-device (SSD300.py:458-462); nothing here has a counterpart there.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

struct odt_ctx {
  void* comm;
  int rank, world;
};

namespace {

struct NcclId {
  char internal[128];
};
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(void**, int, NcclId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*BroadcastFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*ErrStrFn)(int);

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllGatherFn all_gather = nullptr;
  BroadcastFn broadcast = nullptr;
  ErrStrFn err_str = nullptr;
} g_nccl;

constexpr int kNcclInt8 = 0, kNcclFloat32 = 7;  // ncclDataType_t (stable across NCCL 2.x)

int resolve_nccl() {
  if (g_nccl.all_gather) return ODT_OK;
  const char* names[] = {getenv("ODT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    odt::set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    return ODT_ERR_UNSUPPORTED;
  }
  g_nccl.handle = h;
  g_nccl.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  g_nccl.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  g_nccl.broadcast = (BroadcastFn)dlsym(h, "ncclBroadcast");
  g_nccl.err_str = (ErrStrFn)dlsym(h, "ncclGetErrorString");
  AllGatherFn ag = (AllGatherFn)dlsym(h, "ncclAllGather");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.broadcast || !ag) {
    odt::set_error("NCCL library lacks an expected entry point");
    return ODT_ERR_UNSUPPORTED;
  }
  g_nccl.all_gather = ag;
  return ODT_OK;
}

int nccl_ok(int rc, const char* what) {
  if (rc == 0) return ODT_OK;
  odt::set_error("%s failed: NCCL error %d (%s)", what, rc, g_nccl.err_str ? g_nccl.err_str(rc) : "?");
  return ODT_ERR_CUDA;
}

}  // namespace

extern "C" int odt_ctx_unique_id(void* id128_host) {
  ODT_CHECK_ARG(id128_host != nullptr, "null id buffer");
  int rc = resolve_nccl();
  if (rc) return rc;
  NcclId id;
  memset(&id, 0, sizeof(id));
  rc = nccl_ok(g_nccl.get_unique_id(&id), "ncclGetUniqueId");
  if (rc) return rc;
  memcpy(id128_host, &id, sizeof(id));
  return ODT_OK;
}

extern "C" int odt_ctx_create(odt_ctx** ctx, int rank, int world, const void* id128_host) {
  ODT_CHECK_ARG(ctx && id128_host && world >= 1 && rank >= 0 && rank < world, "rank / world / id");
  int rc = resolve_nccl();
  if (rc) return rc;
  NcclId id;
  memcpy(&id, id128_host, sizeof(id));
  void* comm = nullptr;
  rc = nccl_ok(g_nccl.comm_init_rank(&comm, world, id, rank), "ncclCommInitRank");  // on the current CUDA device
  if (rc) return rc;
  odt_ctx* c = (odt_ctx*)malloc(sizeof(odt_ctx));
  if (!c) {
    odt::set_error("out of host memory");
    return ODT_ERR_INVALID;
  }
  c->comm = comm;
  c->rank = rank;
  c->world = world;
  *ctx = c;
  return ODT_OK;
}

extern "C" int odt_ctx_destroy(odt_ctx* ctx) {
  if (!ctx) return ODT_OK;
  int rc = ODT_OK;
  if (ctx->comm && g_nccl.comm_destroy) rc = nccl_ok(g_nccl.comm_destroy(ctx->comm), "ncclCommDestroy");
  free(ctx);
  return rc;
}

extern "C" int odt_allgather_dets(odt_ctx* ctx, const float* rec_local, float* rec_all, long long floats_per_rank,
                                  void* stream) {
  ODT_CHECK_ARG(ctx && ctx->comm && rec_local && rec_all && floats_per_rank > 0, "ctx / buffers / size");
  return nccl_ok(g_nccl.all_gather(rec_local, rec_all, (size_t)floats_per_rank, kNcclFloat32, ctx->comm,
                                   (cudaStream_t)stream),
                 "ncclAllGather");
}

extern "C" int odt_bcast_weights(odt_ctx* ctx, void* buf, long long bytes, int root, void* stream) {
  ODT_CHECK_ARG(ctx && ctx->comm && buf && bytes > 0 && root >= 0 && root < ctx->world, "ctx / buffer / root");
  return nccl_ok(g_nccl.broadcast(buf, buf, (size_t)bytes, kNcclInt8, root, ctx->comm, (cudaStream_t)stream),
                 "ncclBroadcast");
}
