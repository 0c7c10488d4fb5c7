// Gradient exchange of the data-parallel training step: in-place sum all-reduce of one gradient bucket over NCCL
// (NVLink 5 / NVSwitch), issued on the caller's stream so it is ordered by CUDA events against the weight-gradient
// kernels that fill the bucket and can be captured into the training step's CUDA graph.
//
// Replaces the reference's nn.DataParallel gradient path (/root/reference train.py:65-71: replicate / scatter / gather and
// reduce_add_coalesced of the replica gradients onto GPU 0).  One process per GPU; the communicator belongs to this
// library (not to torch.distributed), so its lifetime is under the caller's control: destroy the CUDA graphs that
// captured collectives, then the communicator.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the hosting process already loaded -- PyTorch ships one -- or
// the system library): the build needs no NCCL headers and inference-only users need no NCCL at all.
#include "yb_common.h"
#include <dlfcn.h>
#include <stdlib.h>

namespace yb {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*BroadcastFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

static struct {
  void* handle;
  GetUniqueIdFn get_unique_id;
  CommInitRankFn comm_init_rank;
  AllReduceFn all_reduce;
  BroadcastFn broadcast;
  CommDestroyFn comm_destroy;
  GetErrorStringFn get_error_string;
  GetVersionFn get_version;
} g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

static int nccl_load() {
  if (g_nccl.handle != nullptr) return 0;
  void* h = nullptr;
  const char* forced = getenv("YB_NCCL_PATH");
  if (forced != nullptr && forced[0] != 0) h = dlopen(forced, RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy the process already has (PyTorch's)
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) return fail(YB_ERR_DRIVER, "NCCL not available: %s (set YB_NCCL_PATH)", dlerror());
  g_nccl.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
  g_nccl.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
  g_nccl.broadcast = reinterpret_cast<BroadcastFn>(dlsym(h, "ncclBroadcast"));
  g_nccl.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
  g_nccl.get_error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
  g_nccl.get_version = reinterpret_cast<GetVersionFn>(dlsym(h, "ncclGetVersion"));
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.broadcast || !g_nccl.comm_destroy || !g_nccl.get_error_string)
    return fail(YB_ERR_DRIVER, "NCCL library lacks a required symbol");
  g_nccl.handle = h;
  return 0;
}

static int nccl_check(int rc, const char* what) {
  if (rc == 0) return 0;
  return fail(YB_ERR_DRIVER, "%s: NCCL error %d (%s)", what, rc, g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?");
}

static int nccl_dtype(int dtype, int* out) {
  // yb dtype codes: 0 = float32, 1 = float16, 2 = bfloat16, 3 = int32 -> ncclDataType_t
  static const int map[4] = {7, 6, 9, 2};
  YB_REQUIRE(dtype >= 0 && dtype < 4, "comm: dtype %d (0 f32, 1 f16, 2 bf16, 3 i32)", dtype);
  *out = map[dtype];
  return 0;
}

int comm_version(int* version) {
  int rc = nccl_load();
  if (rc) return rc;
  YB_REQUIRE(version != nullptr, "comm_version: null pointer");
  *version = 0;
  return g_nccl.get_version ? nccl_check(g_nccl.get_version(version), "ncclGetVersion") : 0;
}

int comm_unique_id(void* id128) {
  YB_REQUIRE(id128 != nullptr, "comm_unique_id: null pointer");
  int rc = nccl_load();
  if (rc) return rc;
  return nccl_check(g_nccl.get_unique_id(static_cast<NcclUniqueId*>(id128)), "ncclGetUniqueId");
}

int comm_init(void** comm, int nranks, const void* id128, int rank) {
  YB_REQUIRE(comm != nullptr && id128 != nullptr && nranks > 0 && rank >= 0 && rank < nranks, "comm_init: bad argument");
  int rc = nccl_load();
  if (rc) return rc;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm c = nullptr;
  rc = nccl_check(g_nccl.comm_init_rank(&c, nranks, id, rank), "ncclCommInitRank");
  if (rc) return rc;
  *comm = c;
  return 0;
}

int comm_destroy(void* comm) {
  if (comm == nullptr) return 0;
  int rc = nccl_load();
  if (rc) return rc;
  return nccl_check(g_nccl.comm_destroy(comm), "ncclCommDestroy");
}

int allreduce_bucket(void* comm, void* buf, long long count, int dtype, cudaStream_t stream) {
  YB_REQUIRE(comm != nullptr && buf != nullptr && count > 0, "allreduce_bucket: bad argument");
  int rc = nccl_load();
  if (rc) return rc;
  int dt = 0;
  rc = nccl_dtype(dtype, &dt);
  if (rc) return rc;
  return nccl_check(g_nccl.all_reduce(buf, buf, static_cast<size_t>(count), dt, /*ncclSum*/ 0, comm, stream), "ncclAllReduce");
}

int broadcast_buffer(void* comm, void* buf, long long count, int dtype, int root, cudaStream_t stream) {
  YB_REQUIRE(comm != nullptr && buf != nullptr && count > 0 && root >= 0, "broadcast_buffer: bad argument");
  int rc = nccl_load();
  if (rc) return rc;
  int dt = 0;
  rc = nccl_dtype(dtype, &dt);
  if (rc) return rc;
  return nccl_check(g_nccl.broadcast(buf, buf, static_cast<size_t>(count), dt, root, comm, stream), "ncclBroadcast");
}

}  // namespace yb
