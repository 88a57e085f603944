// Hardware experiment (design input for the weight-gradient kernel):
// tcgen05.mma with BOTH operands MN-major (the contraction index K = pixel is the slow, strided
// index; M = output channel / N = input channel are contiguous, exactly how NHWC activations lie in
// memory).  Questions answered on the device:
//   (1) which descriptor field carries the stride between 64-(or 32-)element blocks along M/N and
//       which the stride between 8-row groups along K for a TMA-written SW128 / SW64 box;
//   (2) may the operand START at a K row that is not a multiple of 8 (the conv's +-1 pixel shift
//       expressed as a start-address offset into a halo box, like the forward row-streaming kernel);
//   (3) does a block stride of 0 replicate a 64-channel box over all 128 M rows (Cout < 128 layers).
// D[128, N] = sum_k A[k][m] * B[k + shift][n] is compared with the CPU result.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
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

constexpr int kKR = 80;    // K rows (pixels) loaded per box
constexpr int kKUse = 64;  // K rows contracted (4 MMAs of K = 16)

struct ExpParams {
  CUtensorMap tmA, tmB;
  int ca, na;       // channels per A box, number of A boxes (M = 128 always)
  int cb, nb;       // channels per B box, number of B boxes (N = cb * nb)
  int shift;        // B starts at K row `shift`
  int swap_fields;  // 0: LBO = block stride, SBO = 8-row group stride; 1: swapped
  int lbo_a_zero;   // replicate A box 0 over all M blocks
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t row_bytes, uint32_t block_stride,
                                            int swap_fields) {
  const uint64_t layout = (row_bytes == 128) ? 2ull : 4ull;
  const uint64_t group = (8ull * row_bytes) >> 4;
  const uint64_t blk = (uint64_t)block_stride >> 4;
  const uint64_t lbo = swap_fields ? group : blk;
  const uint64_t sbo = swap_fields ? blk : group;
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__global__ void __launch_bounds__(128) mn_kernel(const __grid_constant__ ExpParams p, float* __restrict__ out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  const int rowb_a = p.ca * 2, rowb_b = p.cb * 2;
  const int box_a = kKR * rowb_a, box_b = kKR * rowb_b;  // multiples of 1024 (kKR = 80, rows 64/128 B)
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.na * box_a;
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.cb * p.nb;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<128>(&tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, p.na * box_a + p.nb * box_b);
    for (int i = 0; i < p.na; ++i) tma_load_2d(sA + i * box_a, &p.tmA, &full_bar, i * p.ca, 0);
    for (int i = 0; i < p.nb; ++i) tma_load_2d(sB + i * box_b, &p.tmB, &full_bar, i * p.cb, 0);
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, N) | (1u << 15) | (1u << 16);  // A and B MN-major
    for (int j = 0; j < kKUse / 16; ++j) {
      const uint32_t a_addr = smem_u32(sA) + j * 16 * rowb_a;
      const uint32_t b_addr = smem_u32(sB) + (j * 16 + p.shift) * rowb_b;
      const uint64_t da = desc_mn(a_addr, rowb_a, p.lbo_a_zero ? 0 : box_a, p.swap_fields);
      const uint64_t db = desc_mn(b_addr, rowb_b, box_b, p.swap_fields);
      umma_bf16(tmem, da, db, idesc, j != 0);
    }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t acc[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void run(EncodeTiledFn enc, int ca, int cb, int nb, int lbo_a_zero) {
  const int M = 128, N = cb * nb, R = 96;
  std::vector<float> A((size_t)R * M), B((size_t)R * N);
  uint32_t s = 4242;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return __bfloat162float(__float2bfloat16_rn(((s >> 8) & 0xFFFF) / 32768.0f - 1.0f)); };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<__nv_bfloat16> Ab(A.size()), Bb(B.size());
  for (size_t i = 0; i < A.size(); ++i) Ab[i] = __float2bfloat16_rn(A[i]);
  for (size_t i = 0; i < B.size(); ++i) Bb[i] = __float2bfloat16_rn(B[i]);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  CK(cudaMalloc(&dA, Ab.size() * 2));
  CK(cudaMalloc(&dB, Bb.size() * 2));
  CK(cudaMalloc(&dO, (size_t)128 * N * 4));
  CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
  ExpParams p;
  p.ca = ca; p.na = 128 / ca; p.cb = cb; p.nb = nb; p.lbo_a_zero = lbo_a_zero;
  auto mk = [&](CUtensorMap* tm, void* base, int cols, int boxc) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)R};
    cuuint64_t str[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)boxc, (cuuint32_t)kKR};
    cuuint32_t es[2] = {1, 1};
    const CUtensorMapSwizzle sw = boxc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  if (mk(&p.tmA, dA, M, ca) || mk(&p.tmB, dB, N, cb)) { printf("encode failed\n"); return; }
  const size_t smem = (size_t)(128 / ca) * kKR * ca * 2 + (size_t)nb * kKR * cb * 2 + 2048;
  CK(cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // the swapped assignment (LBO = 8-row group, SBO = block) faults (illegal address): first run, r01 log
  for (int swapf = 0; swapf < 1; ++swapf)
    for (int shift = 0; shift <= 9; shift += (shift < 3 ? 1 : 3)) {
      p.swap_fields = swapf;
      p.shift = shift;
      CK(cudaMemset(dO, 0, (size_t)128 * N * 4));
      mn_kernel<<<1, 128, smem>>>(p, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ca=%d cb=%d nb=%d: CUDA error %s\n", ca, cb, nb, cudaGetErrorString(e)); exit(101); }
      std::vector<float> O((size_t)128 * N);
      CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          const int ma = lbo_a_zero ? (m % ca) : m;
          double ref = 0;
          for (int k = 0; k < kKUse; ++k) ref += (double)A[(size_t)k * M + ma] * B[(size_t)(k + shift) * N + n];
          const double err = fabs(ref - O[(size_t)m * N + n]);
          if (err > maxerr) maxerr = err;
          if (err > 1e-2) ++bad;
        }
      printf("[EXP mnmajor] A box %2dch x%d%s  B box %2dch x%d  shift=%d  fields=%s : maxerr=%.4f bad=%d -> %s\n", ca,
             128 / ca, lbo_a_zero ? " (block stride 0)" : "", cb, nb, shift,
             swapf ? "LBO=8-row group, SBO=block" : "LBO=block, SBO=8-row group", maxerr, bad, bad == 0 ? "OK" : "MISMATCH");
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  if (!fp) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  run(enc, 64, 64, 1, 0);
  run(enc, 64, 64, 2, 0);
  run(enc, 32, 32, 2, 0);
  run(enc, 32, 32, 4, 0);
  run(enc, 64, 64, 1, 1);
  run(enc, 32, 64, 1, 1);
  return 0;
}
