// Hardware experiment for the NEXT step: the CTA-pair MMA.  A cluster of two CTAs computes
// D[256 x N] = A[256 x K] * B[N x K]^T with ONE tcgen05.mma.cta_group::2 stream issued by the leader CTA:
//   * each CTA TMA-loads ITS 128 rows of A and ITS N/2 rows of B into its own shared memory (same offsets),
//   * TMEM is allocated with cta_group::2 (one warp per CTA), the two loads are joined by a cluster barrier,
//   * the leader issues M = 256 MMAs (instruction descriptor M = 256) with descriptors built from its own smem
//     addresses, and commits with .multicast::cluster so that BOTH CTAs' barriers are signalled,
//   * each CTA reads its own 128 TMEM lanes = its 128 rows of D.
// This is the building block of a weights-resident row kernel for the 128-channel layers: the pair holds one half of
// the output channels' weights each (147 KB) and every CTA streams only its own pixels.
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


constexpr int kK = 64;  // one 128-byte swizzled row per operand row

// CTA-pair PTX helpers: ptx.cuh (moved there once conv_pair_kernel started using them)

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
    cta2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                float* __restrict__ out /*[256][N]*/, long long* __restrict__ cycles) {
  constexpr int ROWB = kK * 2;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                // this CTA's 128 rows of A
  uint8_t* sB = smem + 128 * ROWB;   // this CTA's N/2 rows of B
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, (128 + N / 2) * ROWB);
    tma_load_2d(sA, &tmA, &full_bar, 0, static_cast<int>(rank) * 128);
    tma_load_2d(sB, &tmB, &full_bar, 0, static_cast<int>(rank) * (N / 2));
    mbar_wait(&full_bar, 0);
  }
  __syncthreads();
  cluster_sync_all();  // both CTAs' operands have landed
  long long t0 = 0;
  if (rank == 0 && threadIdx.x == 0) {
    tc_fence_after();
    const uint64_t da = umma_smem_desc_kmajor(smem_u32(sA), ROWB);
    const uint64_t db = umma_smem_desc_kmajor(smem_u32(sB), ROWB);
    constexpr uint32_t idesc = umma_idesc_bf16(256, N);
    t0 = clock64();
    for (int j = 0; j < kK / 16; ++j) umma_pair_bf16(tmem, da + 2 * j, db + 2 * j, idesc, j != 0);
    for (int it = 0; it < 252; ++it) umma_pair_bf16(tmem + N, da + 2 * (it & 3), db + 2 * (it & 3), idesc, 1u);  // timing only
    umma_pair_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  if (rank == 0 && threadIdx.x == 0) cycles[0] = clock64() - t0;
  tc_fence_after();
  uint32_t acc[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, acc);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)(rank * 128 + warp * 32 + lane) * N + c + i] = __uint_as_float(acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair<512>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int N>
static void run(EncodeTiledFn enc) {
  std::vector<float> A((size_t)256 * kK), B((size_t)N * kK);
  uint32_t s = 4242;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return __bfloat162float(__float2bfloat16_rn(((s >> 8) & 0xFFFF) / 32768.0f - 1.0f)); };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<__nv_bfloat16> Ab(A.size()), Bb(B.size());
  for (size_t i = 0; i < A.size(); ++i) Ab[i] = __float2bfloat16_rn(A[i]);
  for (size_t i = 0; i < B.size(); ++i) Bb[i] = __float2bfloat16_rn(B[i]);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  long long* dC;
  CK(cudaMalloc(&dA, Ab.size() * 2));
  CK(cudaMalloc(&dB, Bb.size() * 2));
  CK(cudaMalloc(&dO, 256 * N * 4));
  CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, 256 * N * 4));
  CUtensorMap tmA, tmB;
  cuuint64_t str[1] = {(cuuint64_t)kK * 2};
  cuuint32_t es[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)kK, 256};
    cuuint32_t box[2] = {(cuuint32_t)kK, 128};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return; }
    cuuint64_t dimsb[2] = {(cuuint64_t)kK, (cuuint64_t)N};
    cuuint32_t boxb[2] = {(cuuint32_t)kK, (cuuint32_t)(N / 2)};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return; }
  }
  const size_t smem = (size_t)(128 + N / 2) * kK * 2 + 2048;
  CK(cudaFuncSetAttribute(cta2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cta2_kernel<N><<<2, 128, smem>>>(tmA, tmB, dO, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[EXP cta2] N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); exit(103); }
  std::vector<float> O(256 * N);
  long long cyc = 0;
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int badrows = 0;
  for (int m = 0; m < 256; ++m) {
    bool bad = false;
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < kK; ++k) ref += (double)A[(size_t)m * kK + k] * B[(size_t)n * kK + k];
      const double err = fabs(ref - O[m * N + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-2) bad = true;
    }
    badrows += bad;
  }
  printf("[EXP cta2] M=256 (2 x 128) N=%3d K=%d: maxerr=%.4f bad_rows=%d -> %s | 256 MMAs in %lld cycles (%.1f per M256 MMA)\n", N, kK,
         maxerr, badrows, badrows == 0 ? "OK" : "MISMATCH", cyc, cyc / 256.0);
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
}


// ---- pipelined K loop with cross-CTA barriers (the skeleton of a 2-CTA conv kernel) ----
// full[st] lives in the LEADER: count 2 (leader's arrive.expect_tx + the peer's remote arrive), both CTAs' TMA loads
// complete_tx on it (the peer through .cta_group::2 with the leader's barrier address); empty[st] lives in every CTA
// and is signalled in both by the leader's multicast tcgen05.commit.
constexpr int kPipeStages = 6;
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
    cta2_pipe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int ksteps,
                     float* __restrict__ out /*[256][N]*/, long long* __restrict__ cyc) {
  constexpr int ROWB = 128;
  constexpr int A_BYTES = 128 * ROWB, B_BYTES = (N / 2) * ROWB, STAGE = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full_bar[kPipeStages], empty_bar[kPipeStages], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kPipeStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const long long t_start = clock64();
  if (warp == 0 && lane == 0) {
    // producer of this CTA
    int st = 0;
    uint32_t ph = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(&empty_bar[st], ph ^ 1);
      const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[st]), 0);
      if (rank == 0) mbar_expect_tx(&full_bar[st], 2 * STAGE);   // arrive + expect both CTAs' bytes
      else mbar_arrive_cluster(leader_full);
      uint8_t* a = smem + st * STAGE;
      tma_load_2d_pair(a, &tmA, leader_full, ks * 64, static_cast<int>(rank) * 128);
      tma_load_2d_pair(a + A_BYTES, &tmB, leader_full, ks * 64, static_cast<int>(rank) * (N / 2));
      if (++st == kPipeStages) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 2 && lane == 0 && rank == 0) {
    // MMA issuer (leader only)
    constexpr uint32_t idesc = umma_idesc_bf16(256, N);
    int st = 0;
    uint32_t ph = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint64_t da = umma_smem_desc_kmajor(smem_u32(smem + st * STAGE), ROWB);
      const uint64_t db = umma_smem_desc_kmajor(smem_u32(smem + st * STAGE + A_BYTES), ROWB);
      for (int j = 0; j < 4; ++j) umma_pair_bf16(tmem, da + 2 * j, db + 2 * j, idesc, (ks | j) != 0);
      umma_pair_commit(&empty_bar[st]);
      if (ks == ksteps - 1) umma_pair_commit(&done_bar);
      if (++st == kPipeStages) {
        st = 0;
        ph ^= 1;
      }
    }
  }
  __syncwarp();
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  if (threadIdx.x == 0 && rank == 0) *cyc = clock64() - t_start;
  uint32_t acc[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, acc);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)(rank * 128 + warp * 32 + lane) * N + c + i] = __uint_as_float(acc[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair<512>(tmem);
}

template <int N>
static void run_pipe(EncodeTiledFn enc, int ksteps) {
  const int K = ksteps * 64;
  std::vector<float> A((size_t)256 * K), B((size_t)N * K);
  uint32_t s = 991;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return __bfloat162float(__float2bfloat16_rn((((s >> 8) & 0xFFFF) / 32768.0f - 1.0f) * 0.25f)); };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<__nv_bfloat16> Ab(A.size()), Bb(B.size());
  for (size_t i = 0; i < A.size(); ++i) Ab[i] = __float2bfloat16_rn(A[i]);
  for (size_t i = 0; i < B.size(); ++i) Bb[i] = __float2bfloat16_rn(B[i]);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  CK(cudaMalloc(&dA, Ab.size() * 2));
  CK(cudaMalloc(&dB, Bb.size() * 2));
  CK(cudaMalloc(&dO, 256 * N * 4));
  CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, 256 * N * 4));
  CUtensorMap tmA, tmB;
  cuuint64_t str[1] = {(cuuint64_t)K * 2};
  cuuint32_t es[2] = {1, 1};
  cuuint64_t dims[2] = {(cuuint64_t)K, 256};
  cuuint32_t box[2] = {64, 128};
  CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode A failed %d\n", (int)r); return; }
  cuuint64_t dimsb[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint32_t boxb[2] = {64, (cuuint32_t)(N / 2)};
  r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode B failed %d\n", (int)r); return; }
  const size_t smem = (size_t)kPipeStages * (128 + N / 2) * 128 + 2048;
  CK(cudaFuncSetAttribute(cta2_pipe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* dC;
  CK(cudaMalloc(&dC, 8));
  cta2_pipe_kernel<N><<<2, 128, smem>>>(tmA, tmB, ksteps, dO, dC);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  cudaFree(dC);
  if (e != cudaSuccess) { printf("[EXP cta2-pipe] N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); exit(104); }
  std::vector<float> O(256 * N);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int badrows = 0;
  for (int m = 0; m < 256; ++m) {
    bool bad = false;
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      const double err = fabs(ref - O[m * N + n]);
      if (err > maxerr) maxerr = err;
      if (err > 2e-2) bad = true;
    }
    badrows += bad;
  }
  printf("[EXP cta2-pipe] M=256 N=%3d K=%4d (%d k-steps, %d-stage ring, peer TMA -> leader barrier, multicast commits): maxerr=%.4f bad_rows=%d -> %s | %lld cycles, %.0f per k-step\n",
         N, K, ksteps, kPipeStages, maxerr, badrows, badrows == 0 ? "OK" : "MISMATCH", cyc, (double)cyc / ksteps);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (!p) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  run<128>(reinterpret_cast<EncodeTiledFn>(p));
  run<64>(reinterpret_cast<EncodeTiledFn>(p));
  run<256>(reinterpret_cast<EncodeTiledFn>(p));
  if (getenv("VUB_EXP_PIPE")) {
    run_pipe<128>(reinterpret_cast<EncodeTiledFn>(p), 18);
    run_pipe<128>(reinterpret_cast<EncodeTiledFn>(p), 5);
    run_pipe<128>(reinterpret_cast<EncodeTiledFn>(p), 180);
    run_pipe<256>(reinterpret_cast<EncodeTiledFn>(p), 180);
  }
  return 0;
}
