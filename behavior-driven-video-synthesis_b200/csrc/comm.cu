// Gradient-bucket exchange of data-parallel training behind the C ABI (`allreduce_bucket(flat,
// ncclComm, stream)`): one in-place all-reduce (average) of a flat fp32 bucket over NVLink, enqueued on the caller's
// stream so that it overlaps the remaining backward launches.  Replaces the gather / reduce half of
// `nn.DataParallel` (experiments/shape_and_pose_net.py:213-214).
//
// NCCL is bound at RUN TIME (dlopen / dlsym): inside a PyTorch process the library torch already loaded is reused
// (one NCCL per process), a stand-alone C++ host gets the system libnccl.so.2; nothing is linked at build time.
#include <dlfcn.h>

#include <mutex>

#include "common.h"

namespace vub {
namespace {

struct NcclUniqueId {
  char internal[128];  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
};
using ncclComm_t = void*;
// nccl.h enums: ncclFloat32 = 7, ncclSum = 0, ncclAvg = 4; ncclSuccess = 0
constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4;

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy this process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
  });
  return a;
}

int nccl_fail(const char* what, int rc) {
  NcclApi& a = api();
  return fail(VUB_E_COMM, "%s: NCCL error %d (%s)", what, rc, a.GetErrorString ? a.GetErrorString(rc) : "?");
}

}  // namespace
}  // namespace vub

using namespace vub;

extern "C" int32_t vub_nccl_version(void) {
  NcclApi& a = api();
  int v = 0;
  if (!a.ok || !a.GetVersion || a.GetVersion(&v) != 0) return 0;
  return v;
}

extern "C" int vub_nccl_unique_id(void* id128) {
  clear_error();
  if (!id128) return fail(VUB_E_NULL, "nccl_unique_id: NULL");
  NcclApi& a = api();
  if (!a.ok) return fail(VUB_E_COMM, "nccl_unique_id: libnccl.so.2 not found");
  NcclUniqueId id;
  const int rc = a.GetUniqueId(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int vub_nccl_comm_init(void** comm, int32_t n_ranks, const void* id128, int32_t rank) {
  clear_error();
  if (!comm || !id128) return fail(VUB_E_NULL, "nccl_comm_init: NULL");
  if (n_ranks <= 0 || rank < 0 || rank >= n_ranks) return fail(VUB_E_SHAPE, "nccl_comm_init: rank %d of %d", rank, n_ranks);
  NcclApi& a = api();
  if (!a.ok) return fail(VUB_E_COMM, "nccl_comm_init: libnccl.so.2 not found");
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  const int rc = a.CommInitRank(&c, n_ranks, id, rank);
  if (rc) return nccl_fail("ncclCommInitRank", rc);
  *comm = c;
  return 0;
}

extern "C" int vub_nccl_comm_destroy(void* comm) {
  clear_error();
  if (!comm) return 0;
  NcclApi& a = api();
  if (!a.ok) return fail(VUB_E_COMM, "nccl_comm_destroy: libnccl.so.2 not found");
  const int rc = a.CommDestroy(comm);
  return rc ? nccl_fail("ncclCommDestroy", rc) : 0;
}

extern "C" int vub_allreduce_bucket(void* flat, int64_t n_floats, void* nccl_comm, int32_t average, void* stream) {
  clear_error();
  if (!flat || !nccl_comm) return fail(VUB_E_NULL, "allreduce_bucket: NULL bucket or communicator");
  if (n_floats <= 0) return fail(VUB_E_SHAPE, "allreduce_bucket: %lld elements", static_cast<long long>(n_floats));
  if (!aligned16(flat)) return fail(VUB_E_ALIGN, "allreduce_bucket: bucket not 16B aligned");
  NcclApi& a = api();
  if (!a.ok) return fail(VUB_E_COMM, "allreduce_bucket: libnccl.so.2 not found");
  const int rc = a.AllReduce(flat, flat, static_cast<size_t>(n_floats), kNcclFloat32, average ? kNcclAvg : kNcclSum,
                             nccl_comm, static_cast<cudaStream_t>(stream));
  return rc ? nccl_fail("ncclAllReduce", rc) : 0;
}
