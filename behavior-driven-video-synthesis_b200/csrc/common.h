// Host-side helpers shared by the C-ABI launchers: error reporting and TMA descriptor encoding.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/vunet_b200.h"

namespace vub {

// Records a message retrievable through vub_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);
void clear_error();

// Returns 0 when the pending launch error state is clean, else the cudaError_t (and records it).
int check_launch(const char* what);

// Launches `kernel(params)` with the programmatic-stream-serialization attribute (programmatic dependent launch):
// the kernel may begin its prologue while its predecessor in the stream is still running and MUST execute
// `pdl_wait()` (ptx.cuh) before it touches anything a predecessor wrote.  VUB_PDL=0 falls back to plain launches.
// `cluster_z` > 1 additionally launches thread-block clusters of (1, 1, cluster_z) CTAs (grid.z a multiple of it).
template <class Params>
inline int launch_pdl(void (*kernel)(Params), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Params& kp,
                      const char* what, int cluster_z = 1) {
  static const bool pdl = !(getenv("VUB_PDL") && getenv("VUB_PDL")[0] == '0');
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cluster_z > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 1;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = static_cast<unsigned>(cluster_z);
    cfg.numAttrs = 2;
  }
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, kp);
  if (le != cudaSuccess) return fail(static_cast<int>(le), "%s launch: %s", what, cudaGetErrorString(le));
  return check_launch(what);
}

// Number of SMs of the current device (queried once per process; 148 on a B200).  Every grid-size / kernel-choice
// threshold uses this, never a literal.
int sm_count();

// Integer environment switch (diagnostics / A-B measurements), read at every call.
inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != '\0') ? atoi(e) : dflt;
}

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Encodes a tiled bf16 tensor map of rank 2..5 (dims/strides innermost first, strides in bytes for
// dims 1..rank-1).  swizzle_bytes is 64 or 128 and must equal box[0]*2.  Returns 0 / error code.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes,
                     bool weights);

}  // namespace vub
