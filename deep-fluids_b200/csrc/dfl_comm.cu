// deepfluids_b200 -- thin NCCL wrapper behind the C-ABI: the ONE exchange step of the data-parallel train step.
//
// batches shard over the GPUs of a box, every rank holds a full replica, and the only collective of a step
// is one ncclAllReduce(sum, fp32) over the flat gradient buffer (NVLink / NVSwitch; NVLS in-switch reduction when NCCL
// selects it).  The reference is single-GPU, so there is no reference interface to mirror: these entry points exist so that
// a host that is NOT PyTorch can run the exchange on the stream the kernels use -- and so that the
// all-reduce can sit INSIDE the captured CUDA graph of the step (NCCL launches are stream-capturable), which
// torch.distributed's own all_reduce call site outside the graph cannot.
// NCCL is resolved at run time (dlopen of libnccl.so.2, preferring the copy already mapped into the process, e.g. the one
// PyTorch bundles): the library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <string.h>

#include "dfl_common.cuh"

namespace dfl {

typedef int ncclResult_like;
struct Id128 { char b[128]; };          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed BY VALUE to ncclCommInitRank
struct NcclApi {
  ncclResult_like (*GetUniqueId)(void*);
  ncclResult_like (*CommInitRank)(void**, int, Id128, int);
  ncclResult_like (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  ncclResult_like (*CommDestroy)(void*);
  const char* (*GetErrorString)(ncclResult_like);
  bool ok;
};
static NcclApi g_nccl = {};

static int nccl_load() {
  if (g_nccl.ok) return DFL_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);     // the copy the host process already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_last_error("dfl_comm: libnccl.so.2 not found (%s)", dlerror());
    return DFL_ERR_INIT;
  }
  *reinterpret_cast<void**>(&g_nccl.GetUniqueId) = dlsym(h, "ncclGetUniqueId");
  *reinterpret_cast<void**>(&g_nccl.CommInitRank) = dlsym(h, "ncclCommInitRank");
  *reinterpret_cast<void**>(&g_nccl.AllReduce) = dlsym(h, "ncclAllReduce");
  *reinterpret_cast<void**>(&g_nccl.CommDestroy) = dlsym(h, "ncclCommDestroy");
  *reinterpret_cast<void**>(&g_nccl.GetErrorString) = dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString) {
    set_last_error("dfl_comm: libnccl.so.2 lacks an expected symbol");
    return DFL_ERR_INIT;
  }
  g_nccl.ok = true;
  return DFL_OK;
}

#define DFL_NCCL_OK(expr, what)                                                          \
  do {                                                                                   \
    const ncclResult_like r__ = (expr);                                                  \
    if (r__ != 0) {                                                                      \
      set_last_error("NCCL error in %s: %s", what, g_nccl.GetErrorString(r__));           \
      return DFL_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

int comm_unique_id(void* id128) {
  DFL_REQUIRE(id128 != nullptr, "dfl_comm_unique_id: null buffer");
  if (int rc = nccl_load()) return rc;
  DFL_NCCL_OK(g_nccl.GetUniqueId(id128), "ncclGetUniqueId");
  return DFL_OK;
}

int comm_init(void** comm, int nranks, const void* id128, int rank) {
  DFL_REQUIRE(comm && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "dfl_comm_init: bad arguments");
  if (int rc = nccl_load()) return rc;
  Id128 id;
  memcpy(id.b, id128, 128);
  DFL_NCCL_OK(g_nccl.CommInitRank(comm, nranks, id, rank), "ncclCommInitRank");
  return DFL_OK;
}

int allreduce(void* buf, size_t count, int dtype, void* comm, cudaStream_t st) {
  DFL_REQUIRE(buf && comm, "dfl_allreduce: null buffer or communicator");
  DFL_REQUIRE(dtype == DT_F32 || dtype == DT_BF16, "dfl_allreduce: dtype must be DFL_F32 or DFL_BF16");
  if (int rc = nccl_load()) return rc;
  const int nccl_dtype = dtype == DT_F32 ? 7 /* ncclFloat32 */ : 9 /* ncclBfloat16 */;
  DFL_NCCL_OK(g_nccl.AllReduce(buf, buf, count, nccl_dtype, 0 /* ncclSum */, comm, st), "ncclAllReduce");
  return DFL_OK;
}

int comm_destroy(void* comm) {
  if (!comm) return DFL_OK;
  if (int rc = nccl_load()) return rc;
  DFL_NCCL_OK(g_nccl.CommDestroy(comm), "ncclCommDestroy");
  return DFL_OK;
}

}  // namespace dfl
