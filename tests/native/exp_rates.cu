// Hardware experiments that bound the conv kernels' two supply chains on a B200 (all 148 SMs busy):
//   1. tcgen05.mma issue/execute rate for the tile shapes the VUnet uses (M=128, N in {32,64,128,256},
//      K=16 per instruction), operands resident in shared memory, accumulating into ONE TMEM tile (the
//      k-loop dependency chain) or alternating between two.
//   2. TMA (cp.async.bulk.tensor) throughput per SM: boxes of 128 rows x 64/128 B streamed through a ring
//      of shared-memory stages with nothing consuming them, from an L2-resident and from an HBM-sized
//      tensor.
// Results are recorded in profiles/r01_exp_rates.log and discussed in DESIGN.md.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
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

// ---------------------------------------------------------------------------------------------
template <int BN, int BKC, int NACC>
__global__ void __launch_bounds__(128) mma_rate_kernel(int n_mma, long long* cycles_out) {
  constexpr int ROWB = BKC * 2;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 + BN) * ROWB / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
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
    const uint32_t hi = umma_desc_hi_kmajor(ROWB);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem));
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 128 * ROWB));
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    const uint32_t issue = lane == 0 ? 1u : 0u;
    const long long t0 = clock64();
    (void)issue;
    for (int i = 0; i < n_mma; i += 8) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t j = u % (BKC / 16);
          const uint32_t d = tmem + (u % NACC) * BN;
          umma_bf16(d, umma_desc_join(hi, a_lo + 2 * j), umma_desc_join(hi, b_lo + 2 * j), idesc, i > 0 ? 1u : 0u);
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

template <int BN, int BKC, int NACC>
static void run_mma() {
  const int n_acc = NACC;
  const int n_mma = 4096;
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  const size_t smem = (128 + BN) * BKC * 2 + 2048;
  CK(cudaFuncSetAttribute(mma_rate_kernel<BN, BKC, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<BN, BKC, NACC><<<148, 128, smem>>>(n_mma, d);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long v : h) avg += v;
  avg /= 148;
  const double per = avg / n_mma;
  const double ideal = 128.0 * BN / 256.0;
  printf("[EXP mma] M=128 N=%3d K=16 row_bytes=%3d accumulators=%d : %.1f cycles/MMA (math floor %.0f) -> %.0f%% of tensor peak\n", BN,
         BKC * 2, n_acc, per, ideal, 100.0 * ideal / per);
  cudaFree(d);
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64) tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int box_bytes, int stages,
                                                      int n_loads, int rows_total, int box_rows, long long* cycles_out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n_boxes = rows_total / box_rows;
    int box = blockIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < n_loads + stages; ++i) {
      const int st = i % stages;
      if (i >= stages) mbar_wait(&full[st], ((i / stages) - 1) & 1);
      if (i < n_loads) {
        mbar_expect_tx(&full[st], box_bytes);
        tma_load_2d(smem + st * box_bytes, &tm, &full[st], 0, (box % n_boxes) * box_rows);
        box += gridDim.x;
      }
    }
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
}

// Same ring, but every load is ONE 1-D bulk copy (cp.async.bulk, no tensor map): what a weight tile costs when it is
// stored in global memory as the exact (pre-swizzled) shared-memory image.  With `mixed` every other load is a
// 128-row x 128 B tensor box (the activation tile of a k-step), the others are bulk copies of the same size.
// bulk_load_1d: csrc/ptx.cuh

__global__ void __launch_bounds__(64) bulk_rate_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* src,
                                                       int box_bytes, int stages, int n_loads, long long n_boxes,
                                                       int mixed, long long* cycles_out) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // n_boxes is a power of two: no divisions in the issuing thread's loop
    uint32_t box = blockIdx.x;
    const uint32_t mask = (uint32_t)n_boxes - 1;
    int st = 0;
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n_loads + stages; ++i) {
      if (i >= stages) mbar_wait(&full[st], ph ^ 1);
      if (i < n_loads) {
        mbar_expect_tx(&full[st], box_bytes);
        const uint32_t b = box & mask;
        if (mixed && (i & 1)) tma_load_2d(smem + st * box_bytes, &tm, &full[st], 0, (int)(b * 128));
        else bulk_load_1d(smem + st * box_bytes, src + (size_t)b * box_bytes, box_bytes, &full[st]);
        box += gridDim.x;
      }
      if (++st == stages) {
        st = 0;
        ph ^= 1;
      }
    }
    cycles_out[blockIdx.x] = clock64() - t0;
  }
}

static void run_bulk(EncodeTiledFn enc, int box_bytes, int stages, size_t tensor_bytes, int mixed, const char* what) {
  void* dA;
  CK(cudaMalloc(&dA, tensor_bytes));
  CK(cudaMemset(dA, 0, tensor_bytes));
  CUtensorMap tm;
  const int rows_total = (int)(tensor_bytes / 128);
  cuuint64_t dims[2] = {64, (cuuint64_t)rows_total};
  cuuint64_t str[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); return; }
  const int n_loads = 2048;
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  const size_t smem = (size_t)stages * box_bytes + 2048;
  CK(cudaFuncSetAttribute(bulk_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_boxes = (long long)(tensor_bytes / box_bytes);
  bulk_rate_kernel<<<148, 64, smem>>>(tm, (const uint8_t*)dA, box_bytes, stages, n_loads, n_boxes, mixed, d);
  bulk_rate_kernel<<<148, 64, smem>>>(tm, (const uint8_t*)dA, box_bytes, stages, n_loads, n_boxes, mixed, d);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long v : h) avg += v;
  avg /= 148;
  printf("[EXP bulk] %-12s %s copy=%5d B stages=%2d : %.1f B/cycle/SM, %.0f cycles/copy\n", what,
         mixed ? "alternating 1-D bulk / 128x128B box" : "1-D bulk only", box_bytes, stages,
         (double)n_loads * box_bytes / avg, avg / n_loads);
  cudaFree(dA); cudaFree(d);
}

static void run_tma(EncodeTiledFn enc, int row_bytes, int box_rows, int stages, size_t tensor_bytes, const char* what) {
  const int rows_total = (int)(tensor_bytes / row_bytes);
  void* dA;
  CK(cudaMalloc(&dA, tensor_bytes));
  CK(cudaMemset(dA, 0, tensor_bytes));
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)row_bytes / 2, (cuuint64_t)rows_total};
  cuuint64_t str[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)row_bytes / 2, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); return; }
  const int box_bytes = row_bytes * box_rows;
  const int n_loads = 2048;
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  const size_t smem = (size_t)stages * box_bytes + 2048;
  CK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  tma_rate_kernel<<<148, 64, smem>>>(tm, box_bytes, stages, n_loads, rows_total, box_rows, d);
  CK(cudaEventRecord(e0));
  tma_rate_kernel<<<148, 64, smem>>>(tm, box_bytes, stages, n_loads, rows_total, box_rows, d);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long v : h) avg += v;
  avg /= 148;
  const double bytes = (double)n_loads * box_bytes;
  printf("[EXP tma] %-12s row_bytes=%3d box=%3d rows (%5d B) stages=%2d : %.1f B/cycle/SM, %.0f cycles/box, chip %.2f TB/s\n", what,
         row_bytes, box_rows, box_bytes, stages, bytes / avg, avg / n_loads, bytes * 148 / (ms * 1e-3) / 1e12);
  cudaFree(dA); cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "bulk")) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
    for (int stages : {4, 8}) {
      run_bulk(enc, 16384, stages, 32ull << 20, 0, "L2-resident");
      run_bulk(enc, 16384, stages, 32ull << 20, 1, "L2-resident");
      run_bulk(enc, 8192, stages, 32ull << 20, 0, "L2-resident");
      run_bulk(enc, 32768, stages > 6 ? 6 : stages, 32ull << 20, 0, "L2-resident");
    }
    run_bulk(enc, 16384, 8, 4ull << 30, 0, "HBM");
    return 0;
  }
  run_mma<16, 32, 1>();
  run_mma<32, 32, 1>();
  run_mma<32, 32, 2>();
  run_mma<32, 64, 1>();
  run_mma<64, 64, 1>();
  run_mma<64, 64, 2>();
  run_mma<128, 64, 1>();
  run_mma<128, 64, 2>();
  run_mma<256, 64, 1>();
  run_mma<256, 64, 2>();
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  for (int stages : {2, 4, 8}) {
    run_tma(enc, 128, 128, stages, 32ull << 20, "L2-resident");
    run_tma(enc, 128, 128, stages, 4ull << 30, "HBM");
    run_tma(enc, 64, 128, stages, 32ull << 20, "L2-resident");
    run_tma(enc, 64, 128, stages, 4ull << 30, "HBM");
  }
  run_tma(enc, 128, 256, 4, 4ull << 30, "HBM");
  run_tma(enc, 128, 64, 8, 4ull << 30, "HBM");
  return 0;
}
