// Hardware experiment (design input for the row-streaming conv kernel):
// can a K-major swizzled UMMA operand START at a row that is not a multiple of 8 rows (i.e. not
// aligned to the swizzle atom), and if so does the descriptor's base_offset field have to carry
// (addr >> 7) & 7?  A [R x 64] (SW128) or [R x 32] (SW64) bf16 matrix is TMA-loaded into shared
// memory once; D = A[shift : shift+128, :] * B^T is computed for shift = 0..9 with both descriptor
// variants and compared with the CPU result.
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

constexpr int kRows = 144;  // rows loaded into smem (>= 128 + max shift, multiple of 8)
constexpr int kN = 32;

template <int BKC>
__global__ void __launch_bounds__(128)
    rowshift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    int shift, int use_base_offset, float* __restrict__ out /*[128][kN]*/) {
  constexpr int ROWB = BKC * 2;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kRows * ROWB;  // kRows*ROWB is a multiple of 1024
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<32>(&tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, (kRows + kN) * ROWB);
    tma_load_2d(sA, &tmA, &full_bar, 0, 0);
    tma_load_2d(sB, &tmB, &full_bar, 0, 0);
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + shift * ROWB;
    uint64_t da = umma_smem_desc_kmajor(a_addr, ROWB);
    if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t db = umma_smem_desc_kmajor(smem_u32(sB), ROWB);
    constexpr uint32_t idesc = umma_idesc_bf16(128, kN);
    for (int j = 0; j < BKC / 16; ++j) umma_bf16(tmem, da + 2 * j, db + 2 * j, idesc, j != 0);
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  uint32_t acc[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), acc);
  tmem_ld_wait();
  for (int i = 0; i < kN; ++i) out[(warp * 32 + lane) * kN + i] = __uint_as_float(acc[i]);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<32>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BKC>
static void run(EncodeTiledFn enc) {
  const int R = 160;
  std::vector<float> A((size_t)R * BKC), B((size_t)kN * BKC);
  uint32_t s = 777;
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
  CK(cudaMalloc(&dO, 128 * kN * 4));
  CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  const CUtensorMapSwizzle sw = BKC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    cuuint64_t dims[2] = {(cuuint64_t)BKC, (cuuint64_t)R};
    cuuint64_t str[1] = {(cuuint64_t)BKC * 2};
    cuuint32_t box[2] = {(cuuint32_t)BKC, (cuuint32_t)kRows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return; }
    cuuint64_t dimsb[2] = {(cuuint64_t)BKC, (cuuint64_t)kN};
    cuuint32_t boxb[2] = {(cuuint32_t)BKC, (cuuint32_t)kN};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return; }
  }
  const size_t smem = (size_t)(kRows + kN) * BKC * 2 + 2048;
  CK(cudaFuncSetAttribute(rowshift_kernel<BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int shift = 0; shift <= 9; ++shift)
    for (int ubo = 0; ubo < 2; ++ubo) {
      CK(cudaMemset(dO, 0, 128 * kN * 4));
      rowshift_kernel<BKC><<<1, 128, smem>>>(tmA, tmB, shift, ubo, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("BKC=%d shift=%d bo=%d: CUDA error %s\n", BKC, shift, ubo, cudaGetErrorString(e)); exit(101); }
      std::vector<float> O(128 * kN);
      CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      int badrows = 0;
      for (int m = 0; m < 128; ++m) {
        bool bad = false;
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < BKC; ++k) ref += (double)A[(size_t)(m + shift) * BKC + k] * B[(size_t)n * BKC + k];
          const double err = fabs(ref - O[m * kN + n]);
          if (err > maxerr) maxerr = err;
          if (err > 1e-2) bad = true;
        }
        badrows += bad;
      }
      printf("[EXP rowshift] row_bytes=%3d shift=%d base_offset_field=%s : maxerr=%.4f bad_rows=%d -> %s\n", BKC * 2, shift,
             ubo ? "(addr>>7)&7" : "0", maxerr, badrows, badrows == 0 ? "OK" : "MISMATCH");
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (!p) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  run<64>(reinterpret_cast<EncodeTiledFn>(p));
  run<32>(reinterpret_cast<EncodeTiledFn>(p));
  return 0;
}
