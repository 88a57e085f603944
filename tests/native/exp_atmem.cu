// Hardware experiment (design input for the next step of the row-streaming conv kernel):
// can the A operand of tcgen05.mma live in TMEM, filled by tcgen05.cp (shared memory -> TMEM) from the SAME
// K-major swizzled tile a TMA load wrote, starting at an arbitrary pixel row (the conv's column tap)?  If so, an
// input row has to be read from shared memory 3 times (one copy per column tap) instead of 9 (once per MMA tap),
// which is what bounds the wide layers (profiles/r01_ncu_rows_smem.txt).
//   1. A [R x BKC] bf16 is TMA-loaded (SW64 / SW128); for every K16 slice j:
//        tcgen05.cp.cta_group::1.128x256b [tA + 8*j], desc(A + shift rows, slice j)
//   2. the TMEM image is read back and compared with the expected packing (lane = row, column c = K pair c);
//   3. D = A_tmem * B^T with tcgen05.mma (A from TMEM) is compared with the CPU result;
//   4. timing: 256 back-to-back MMAs with A from TMEM vs A from shared memory, N = 32 / 64.
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


constexpr int kRows = 144;

// tmem_cp_128x256b / umma_bf16_ts: csrc/ptx.cuh

template <int BKC, int N>
__global__ void __launch_bounds__(128)
    atmem_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift,
                 float* __restrict__ out /*[128][N]*/, uint32_t* __restrict__ a_img /*[128][BKC/2]*/,
                 long long* __restrict__ cycles /*[2]*/) {
  constexpr int ROWB = BKC * 2;
  constexpr uint32_t A_COL = 64;  // TMEM column of the A image (accumulator at column 0)
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kRows * ROWB;
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  constexpr uint32_t idesc = umma_idesc_bf16(128, N);
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, (kRows + N) * ROWB);
    tma_load_2d(sA, &tmA, &full_bar, 0, 0);
    tma_load_2d(sB, &tmB, &full_bar, 0, 0);
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint64_t da = umma_smem_desc_kmajor(smem_u32(sA) + shift * ROWB, ROWB);
    const uint64_t db = umma_smem_desc_kmajor(smem_u32(sB), ROWB);
    for (int j = 0; j < BKC / 16; ++j) tmem_cp_128x256b(tmem + A_COL + 8 * j, da + 2 * j);
    for (int j = 0; j < BKC / 16; ++j) umma_bf16_ts(tmem, tmem + A_COL + 8 * j, db + 2 * j, idesc, j != 0);
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, phase);
  phase ^= 1;
  tc_fence_after();
  {
    uint32_t acc[32];
    for (int c = 0; c < N; c += 32) {
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, acc);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c + i] = __uint_as_float(acc[i]);
    }
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + A_COL, acc);
    tmem_ld_wait();
    for (int i = 0; i < BKC / 2; ++i) a_img[(warp * 32 + lane) * (BKC / 2) + i] = acc[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- timing: 256 MMAs, A from TMEM vs A from shared memory ----
  for (int mode = 0; mode < 2; ++mode) {
    long long t0 = 0;
    if (threadIdx.x == 0) {
      const uint64_t da = umma_smem_desc_kmajor(smem_u32(sA), ROWB);
      const uint64_t db = umma_smem_desc_kmajor(smem_u32(sB), ROWB);
      t0 = clock64();
      for (int it = 0; it < 256; ++it) {
        const int j = it % (BKC / 16);
        if (mode == 0) umma_bf16_ts(tmem, tmem + A_COL + 8 * j, db + 2 * j, idesc, 1u);
        else umma_bf16(tmem, da + 2 * j, db + 2 * j, idesc, 1u);
      }
      umma_commit(&done_bar);
    }
    mbar_wait(&done_bar, phase);
    phase ^= 1;
    if (threadIdx.x == 0) cycles[mode] = clock64() - t0;
    tc_fence_after();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem);
}


// tmem_st16 / tmem_st_wait: csrc/ptx.cuh

// A written to tensor memory by the four warps from REGISTERS (each warp its own lane quarter), then MMA with A from
// TMEM: the path an ELU-on-load transform would use to skip shared memory for the transformed operand.
template <int BKC, int N>
__global__ void __launch_bounds__(128)
    atmem_st_kernel(const __nv_bfloat16* __restrict__ A /*[R][BKC]*/, const __grid_constant__ CUtensorMap tmB, int shift,
                    float* __restrict__ out /*[128][N]*/) {
  constexpr int ROWB = BKC * 2;
  constexpr uint32_t A_COL = 64;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sB = smem;
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    mbar_expect_tx(&full_bar, N * ROWB);
    tma_load_2d(sB, &tmB, &full_bar, 0, 0);
  }
  // every thread: its row of A (row = TMEM lane + shift), 16 channels = 8 packed columns per K16 slice
  const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)(warp * 32 + lane + shift) * BKC);
#pragma unroll
  for (int c0 = 0; c0 < BKC / 2; c0 += 16) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = arow[c0 + i];
    tmem_st16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + A_COL + c0, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint64_t db = umma_smem_desc_kmajor(smem_u32(sB), ROWB);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    for (int j = 0; j < BKC / 16; ++j) umma_bf16_ts(tmem, tmem + A_COL + 8 * j, db + 2 * j, idesc, j != 0);
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  uint32_t acc[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, acc);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c + i] = __uint_as_float(acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BKC, int N>
static void run(EncodeTiledFn enc) {
  const int R = 160;
  std::vector<float> A((size_t)R * BKC), B((size_t)N * BKC);
  uint32_t s = 777;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return __bfloat162float(__float2bfloat16_rn(((s >> 8) & 0xFFFF) / 32768.0f - 1.0f)); };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<__nv_bfloat16> Ab(A.size()), Bb(B.size());
  for (size_t i = 0; i < A.size(); ++i) Ab[i] = __float2bfloat16_rn(A[i]);
  for (size_t i = 0; i < B.size(); ++i) Bb[i] = __float2bfloat16_rn(B[i]);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  uint32_t* dI;
  long long* dC;
  CK(cudaMalloc(&dA, Ab.size() * 2));
  CK(cudaMalloc(&dB, Bb.size() * 2));
  CK(cudaMalloc(&dO, 128 * N * 4));
  CK(cudaMalloc(&dI, 128 * (BKC / 2) * 4));
  CK(cudaMalloc(&dC, 16));
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
    cuuint64_t dimsb[2] = {(cuuint64_t)BKC, (cuuint64_t)N};
    cuuint32_t boxb[2] = {(cuuint32_t)BKC, (cuuint32_t)N};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return; }
  }
  const size_t smem = (size_t)(kRows + N) * BKC * 2 + 2048;
  CK(cudaFuncSetAttribute(atmem_kernel<BKC, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int shift : {0, 1, 2, 7, 9}) {
    CK(cudaMemset(dO, 0, 128 * N * 4));
    CK(cudaMemset(dI, 0, 128 * (BKC / 2) * 4));
    atmem_kernel<BKC, N><<<1, 128, smem>>>(tmA, tmB, shift, dO, dI, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("BKC=%d N=%d shift=%d: CUDA error %s\n", BKC, N, shift, cudaGetErrorString(e)); exit(101); }
    std::vector<float> O(128 * N);
    std::vector<uint32_t> I(128 * (BKC / 2));
    long long cyc[2];
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(I.data(), dI, I.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
    int img_bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int c = 0; c < BKC / 2; ++c) {
        const uint16_t lo = *reinterpret_cast<const uint16_t*>(&Ab[(size_t)(m + shift) * BKC + 2 * c]);
        const uint16_t hi = *reinterpret_cast<const uint16_t*>(&Ab[(size_t)(m + shift) * BKC + 2 * c + 1]);
        if (I[m * (BKC / 2) + c] != (((uint32_t)hi << 16) | lo)) ++img_bad;
      }
    double maxerr = 0;
    int badrows = 0;
    for (int m = 0; m < 128; ++m) {
      bool bad = false;
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < BKC; ++k) ref += (double)A[(size_t)(m + shift) * BKC + k] * B[(size_t)n * BKC + k];
        const double err = fabs(ref - O[m * N + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-2) bad = true;
      }
      badrows += bad;
    }
    printf("[EXP atmem] row_bytes=%3d N=%2d shift=%d : tmem image mismatches=%d  D maxerr=%.4f bad_rows=%d -> %s | 256 MMAs: A in TMEM %lld cyc (%.1f/MMA), A in smem %lld cyc (%.1f/MMA)\n",
           BKC * 2, N, shift, img_bad, maxerr, badrows, badrows == 0 ? "OK" : "MISMATCH", cyc[0], cyc[0] / 256.0, cyc[1], cyc[1] / 256.0);
    if (img_bad && shift == 0) {
      printf("   first rows of the TMEM image (got | expected):\n");
      for (int m = 0; m < 4; ++m) {
        printf("   lane %d:", m);
        for (int c = 0; c < 8; ++c) printf(" %08x", I[m * (BKC / 2) + c]);
        printf(" |");
        for (int c = 0; c < 8; ++c) {
          const uint16_t lo = *reinterpret_cast<const uint16_t*>(&Ab[(size_t)m * BKC + 2 * c]);
          const uint16_t hi = *reinterpret_cast<const uint16_t*>(&Ab[(size_t)m * BKC + 2 * c + 1]);
          printf(" %08x", ((uint32_t)hi << 16) | lo);
        }
        printf("\n");
      }
    }
  }
  CK(cudaFuncSetAttribute(atmem_st_kernel<BKC, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int shift : {0, 1, 2}) {
    CK(cudaMemset(dO, 0, 128 * N * 4));
    atmem_st_kernel<BKC, N><<<1, 128, smem>>>(dA, tmB, shift, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("st BKC=%d N=%d shift=%d: CUDA error %s\n", BKC, N, shift, cudaGetErrorString(e)); exit(102); }
    std::vector<float> O(128 * N);
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int badrows = 0;
    for (int m = 0; m < 128; ++m) {
      bool bad = false;
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < BKC; ++k) ref += (double)A[(size_t)(m + shift) * BKC + k] * B[(size_t)n * BKC + k];
        const double err = fabs(ref - O[m * N + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-2) bad = true;
      }
      badrows += bad;
    }
    printf("[EXP atmem-st] row_bytes=%3d N=%2d shift=%d : A written with tcgen05.st from registers, D maxerr=%.4f bad_rows=%d -> %s\n",
           BKC * 2, N, shift, maxerr, badrows, badrows == 0 ? "OK" : "MISMATCH");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dI); cudaFree(dC);
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (!p) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  run<32, 32>(reinterpret_cast<EncodeTiledFn>(p));
  run<64, 64>(reinterpret_cast<EncodeTiledFn>(p));
  run<64, 32>(reinterpret_cast<EncodeTiledFn>(p));
  return 0;
}
