// Hardware experiment: tcgen05.mma execute rate with MN-major operands (the weight-gradient kernel's
// layout: pixels = K are the slow index), all 148 SMs busy, operands resident in shared memory.
// Variants: channel-block stride (LBO) = box size (separate TMA boxes), 0 (replicated box: layers with
// fewer than 128 output channels) and one pixel row (the three column taps of a halo box as three N blocks).
// Results: profiles/r01_exp_rates_mn.log, discussed in DESIGN.md.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../behavior-driven-video-synthesis_b200/csrc/ptx.cuh"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(100);                                                                      \
    }                                                                                 \
  } while (0)

using namespace vub;

struct Cfg {
  int a_rowb, a_lbo;  // A: row bytes (128 = 64 ch boxes, 64 = 32 ch boxes), block stride in bytes
  int b_rowb, b_lbo;  // B likewise
  int n;              // MMA N
  int a_major_mn, b_major_mn;
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t rowb, uint32_t lbo) {
  const uint64_t layout = (rowb == 128) ? 2ull : 4ull;
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)((8 * rowb) >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}

__global__ void __launch_bounds__(128) mn_rate_kernel(Cfg c, int n_mma, long long* cycles_out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(&tmem_slot);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = umma_idesc_bf16(128, c.n) | (c.a_major_mn ? (1u << 15) : 0u) | (c.b_major_mn ? (1u << 16) : 0u);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          // K slice u: MN-major advances 16 pixel rows, K-major 32 bytes inside the row
          const uint32_t a_off = c.a_major_mn ? u * 16 * c.a_rowb : u * 32;
          const uint32_t b_off = c.b_major_mn ? u * 16 * c.b_rowb : u * 32;
          umma_bf16(tmem, mk_desc(a_base + a_off, c.a_rowb, c.a_lbo), mk_desc(b_base + b_off, c.b_rowb, c.b_lbo), idesc,
                    i > 0 ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    if (lane == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

static void run(const char* what, Cfg c) {
  const int n_mma = 4096;
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  const size_t smem = 161 * 1024 + 1024;
  CK(cudaFuncSetAttribute(mn_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) mn_rate_kernel<<<148, 128, smem>>>(c, n_mma, d);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long v : h) avg += v;
  avg /= 148;
  const double per = avg / n_mma, ideal = 128.0 * c.n / 256.0;
  printf("[EXP mma-mn] %-58s N=%3d : %6.1f cycles/MMA (math floor %3.0f) -> %3.0f%% of tensor peak\n", what, c.n, per, ideal,
         100.0 * ideal / per);
  cudaFree(d);
}

int main() {
  const int box64 = 64 * 128, box32 = 64 * 64;  // 64-pixel boxes
  run("K-major A and B (forward conv reference)", Cfg{128, 16, 128, 16, 128, 0, 0});
  run("MN A 2x64ch boxes, MN B 2x64ch boxes", Cfg{128, box64, 128, box64, 128, 1, 1});
  run("MN A 2x64ch boxes, MN B 4x64ch boxes", Cfg{128, box64, 128, box64, 256, 1, 1});
  run("MN A 2x64ch boxes, MN B 1x64ch box", Cfg{128, box64, 128, box64, 64, 1, 1});
  run("MN A 64ch box replicated (LBO 0), MN B 1x64ch", Cfg{128, 0, 128, box64, 64, 1, 1});
  run("MN A 64ch box replicated (LBO 0), MN B 3x64ch", Cfg{128, 0, 128, box64, 192, 1, 1});
  run("MN A 32ch box replicated x4 (LBO 0), MN B 1x32ch", Cfg{64, 0, 64, box32, 32, 1, 1});
  run("MN A 32ch box replicated x4 (LBO 0), MN B 8x32ch", Cfg{64, 0, 64, box32, 256, 1, 1});
  run("MN A 4x32ch boxes, MN B 8x32ch", Cfg{64, box32, 64, box32, 256, 1, 1});
  run("MN A 2x64ch, MN B halo: 3 blocks x 64ch, LBO = 1 pixel row", Cfg{128, box64, 128, 128, 192, 1, 1});
  run("MN A 32ch x4 (LBO 0), MN B halo: 3 blocks x 32ch, LBO = 1 row", Cfg{64, 0, 64, 64, 96, 1, 1});
  run("MN A 2x64ch, ones block N=16 (LBO 0)", Cfg{128, box64, 128, 0, 16, 1, 1});
  run("K-major A, MN B 2x64ch", Cfg{128, 16, 128, box64, 128, 0, 1});
  run("MN A 2x64ch, K-major B", Cfg{128, box64, 128, 16, 128, 1, 0});
  return 0;
}
