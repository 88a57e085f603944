// Forward weight-normalised convolution as an implicit GEMM on tcgen05 (sm_100a).
//
//   D[128 pixels, BN channels] (TMEM, fp32) += A[128, BKC] (smem, bf16) * B[BN, BKC]^T (smem, bf16)
//
// A k-step is one (source, filter tap, BKC-channel block).  The A tile of a k-step is a 4-D TMA box
// (BKC channels x TW x TH x TN pixels, TW*TH*TN = 128) of the NHWC activation tensor, shifted by the
// tap offset; out-of-bounds coordinates (the conv's zero padding, and batch tails) are zero-filled
// by TMA, so no im2col matrix ever exists.  Stride-2 convs read four parity sub-grids of the input
// through four tensor maps.  TMA writes the box with 64B/128B swizzle, which is exactly the canonical
// K-major UMMA operand layout (one pixel = one row).
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> fp32 scale/shift/residual/ELU -> global).
// One CTA = one 128-pixel tile x one BN tile; several CTAs are resident per SM so one CTA's
// epilogue overlaps its neighbours' main loops.
//
// Replaces: NormConv2d.forward lib/modules.py:140-145 (+ VunetRNB's cat/ELU/add 221-233,
// DepthToSpace 24-34, Downsample 148-161) of the reference.
#include <cuda_bf16.h>

#include <cstring>
#include <mutex>

#include <cstdlib>

#include "common.h"
#include "conv_epilogue.cuh"
#include "conv_slab.cuh"
#include "conv_tiles.cuh"
#include "ptx.cuh"

namespace vub {

constexpr int kTileM = 128;
constexpr int kConvThreads = 192;
constexpr int kMaxStages = 8;

struct ConvKParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  EpiParams ep;
  int batch;
  int nsrc, cblocks, ntaps, stride;
  int tw, th, tn, tw_log2, th_log2;
  int tiles_w, tiles_h;
  int stages;
  const uint8_t* w_tiled;  // tile-major pre-swizzled weights [K/64][w_rows][128 B], or nullptr (BKC == 64 only)
  int w_rows;
  int ksplit;  // split-K instantiation: CTAs per cluster (= gridDim.z), each takes a contiguous range of the k-steps
  int reduce_scatter;  // split-K by reduce-scatter (every CTA owns 32 columns of the tile) instead of a reducing leader
};

// Epilogue for CH (16 or 32) consecutive channels of one pixel held in registers.
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, const uint32_t (&acc)[CH],
                                               const float* s_scale, const float* s_shift, int c_local,
                                               int c_global, bool valid, int n, int oh, int ow) {
  float y[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i)
    y[i] = fmaf(__uint_as_float(acc[i]), s_scale[c_local + i], s_shift[c_local + i]);
  if (!valid) return;
  if (p.residual != nullptr) {
    const long long pix = ((long long)n * p.hout + oh) * p.wout + ow;
    const uint4* r = reinterpret_cast<const uint4*>(p.residual + pix * p.cout + c_global);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j) add_bf16x8(__ldg(r + j), &y[8 * j]);
  }
  epi_store<CH>(p, y, c_global, n, oh, ow);
}

// fp32 partial accumulators of the non-leader CTAs of a split-K cluster, parked in the LEADER's shared memory behind
// the epilogue staging tiles: [peer - 1][column][row] (a warp's 32 rows of one column = 128 contiguous bytes).
template <int BN>
struct SplitKPart {
  static constexpr int GW = BN < 128 ? BN : 128;
  static constexpr int OFF = BN >= 32 ? 4 * EpiStage<(GW < 32 ? 32 : GW)>::WARP_BYTES : 0;
  static constexpr int PEER_BYTES = BN * kTileM * 4;
};

// SPLITK: the k-steps of one output tile are split over the CTAs of a thread-block cluster (cluster dims (1,1,ksplit)).
// Every CTA accumulates its range in its own tensor memory; the peers then push their fp32 partials into the leader's
// shared memory through DSMEM, and the leader adds them (in rank order: deterministic) before the usual epilogue.
// For the smallest maps (<= 8x8), where a handful of CTAs would otherwise each stream an 18 .. 144-step K loop.
//
// SPLITK == 2, reduce-scatter (linear layers: few pixel tiles, thousands of output features): instead of one leader
// collecting (ksplit - 1) full partial tiles -- which caps ksplit at 320 / BN -- every CTA of the cluster OWNS the
// column slice [rank * BN / ksplit, +BN / ksplit) of the tile: it pushes the other slices of its accumulator to their
// owners through DSMEM, keeps its own in tensor memory, and after the second cluster barrier reduces and stores its
// 32 columns itself.  With BN = 256 and eight ranges a CTA streams an eighth of the activation rows and of its weight
// tile, receives 7 x 16 KB, and all eight epilogues run in parallel.
template <int BN, int BKC, int SPLITK>
__global__ void __launch_bounds__(kConvThreads)
    conv_fwd_kernel(const __grid_constant__ ConvKParams p) {
  constexpr int ROWB = BKC * 2;
  constexpr int A_BYTES = kTileM * ROWB;
  constexpr int B_BYTES = BN * ROWB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024B alignment");

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_scale[BN];
  __shared__ float s_shift[BN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  int t = blockIdx.x;
  const int twi = t % p.tiles_w;
  t /= p.tiles_w;
  const int thi = t % p.tiles_h;
  const int tni = t / p.tiles_h;
  const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.tn;
  const int S = p.stages;
  const int KS = p.nsrc * p.ntaps * p.cblocks;
  int ks_begin = 0, ks_end = KS;
  uint32_t crank = 0;
  if constexpr (SPLITK != 0) {
    crank = cluster_ctarank();
    ks_begin = KS * static_cast<int>(crank) / p.ksplit;
    ks_end = KS * (static_cast<int>(crank) + 1) / p.ksplit;
  }

  // This kernel serves the small maps, where a launch is ~10 us of mostly latency and ~45 of them run back to
  // back: let the next kernel's CTAs start their prologue now, and do ours (barriers, TMEM, descriptor fetch)
  // before waiting for the previous kernel's results.
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<TMEM_COLS>(&tmem_base_slot);
    tmem_relinquish();
  }
  pdl_wait();  // everything below may read what the previous kernel(s) wrote (activations, prepared weights)
  for (int i = threadIdx.x; i < BN; i += kConvThreads) {
    s_scale[i] = p.ep.scale[n_tile * BN + i];
    s_shift[i] = p.ep.shift[n_tile * BN + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Producer / MMA loops are executed by all lanes of their warp (warp-uniform control flow); the TMA /
  // tcgen05 instructions are issued from `if (elect_one())` blocks so ptxas emits back-to-back UTCHMMA
  // with uniform-register descriptors instead of a per-lane "waterfall" loop around each one.
  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    int st = 0;
    uint32_t ph = 0;
    for (int ks = ks_begin; ks < ks_end; ++ks) {
      const int cb = ks % p.cblocks;
      const int st_ = ks / p.cblocks;
      const int tap = st_ % p.ntaps;
      const int src = st_ / p.ntaps;
      int dh = 0, dw = 0, mi = src;
      if (p.ntaps == 9) {
        const int r = tap / 3, s = tap - 3 * r;
        if (p.stride == 1) {
          dh = r - 1;
          dw = s - 1;
        } else {  // input row 2*oh + r - 1: parity (r != 1), coordinate oh - (r == 0)
          mi = ((r != 1) ? 2 : 0) + ((s != 1) ? 1 : 0);
          dh = (r == 0) ? -1 : 0;
          dw = (s == 0) ? -1 : 0;
        }
      }
      mbar_wait(&empty_bar[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[st], STAGE_BYTES);
        uint8_t* a = smem + st * STAGE_BYTES;
        tma_load_4d(a, &p.tmA[mi], &full_bar[st], cb * BKC, w0 + dw, h0 + dh, n0);
        if (BKC == 64 && p.w_tiled != nullptr)  // one 1-D bulk copy instead of a BN-row tensor box
          bulk_load_1d(a + A_BYTES, p.w_tiled + ((size_t)ks * p.w_rows + n_tile * BN) * 128, B_BYTES, &full_bar[st]);
        else
          tma_load_2d(a + A_BYTES, &p.tmB, &full_bar[st], ks * BKC, n_tile * BN);
      }
      __syncwarp();
      if (++st == S) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
    const uint32_t desc_hi = umma_desc_hi_kmajor(ROWB);
    const uint32_t smem_lo = umma_desc_lo(smem_u32(smem));
    int st = 0;
    uint32_t ph = 0;
    for (int ks = ks_begin; ks < ks_end; ++ks) {
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint32_t a_lo = smem_lo + st * (STAGE_BYTES >> 4);
      const uint32_t b_lo = a_lo + (A_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < BKC / 16; ++j)
          umma_bf16(tmem_base, umma_desc_join(desc_hi, a_lo + 2 * j), umma_desc_join(desc_hi, b_lo + 2 * j),
                    idesc, (ks != ks_begin || j != 0) ? 1u : 0u);
        umma_commit(&empty_bar[st]);
        if (ks == ks_end - 1) umma_commit(&tmem_full_bar);
      }
      __syncwarp();
      if (++st == S) {
        st = 0;
        ph ^= 1;
      }
    }
  } else {
    mbar_wait(&tmem_full_bar, 0);  // epilogue warps: this CTA's accumulator is complete
    tc_fence_after();
  }

  if constexpr (SPLITK == 2) {
    // ---- reduce-scatter over the cluster: slice width W = BN / ksplit = 32 columns ----
    constexpr int W = 32;
    tc_fence_before();
    cluster_sync_all();  // #1: every CTA's accumulator is complete, every CTA's pipeline stages are idle
    tc_fence_after();
    if (warp >= 2) {
      const int q = warp & 3;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int peer = 0; peer < p.ksplit; ++peer) {
        if (peer == static_cast<int>(crank)) continue;
        const int slot = static_cast<int>(crank) < peer ? static_cast<int>(crank) : static_cast<int>(crank) - 1;
        uint32_t acc[32];
        tmem_ld32(trow + peer * W, acc);
        tmem_ld_wait();
        const uint32_t dst = mapa_u32(smem_u32(smem) + slot * (W * kTileM * 4) + (q * 32 + lane) * 4, peer);
#pragma unroll
        for (int i = 0; i < 32; ++i) st_cluster_u32(dst + i * (kTileM * 4), acc[i]);
      }
    }
    tc_fence_before();
    cluster_sync_all();  // #2: this CTA's slice has all its partials
    tc_fence_after();
    if (warp >= 2) {
      const int q = warp & 3;
      const int m = q * 32 + lane;
      const int dw = m & (p.tw - 1);
      const int dh = (m >> p.tw_log2) & (p.th - 1);
      const int dn = m >> (p.tw_log2 + p.th_log2);
      const int n = n0 + dn, oh = h0 + dh, ow = w0 + dw;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t acc[32];
      tmem_ld32(trow + crank * W, acc);
      tmem_ld_wait();
      const uint32_t src = smem_u32(smem) + (q * 32 + lane) * 4;
      for (int slot = 0; slot < p.ksplit - 1; ++slot) {  // ranks in increasing order, this CTA's own range at its place
#pragma unroll
        for (int i = 0; i < 32; ++i)
          acc[i] = __float_as_uint(__uint_as_float(acc[i]) + lds_f32(src + slot * (W * kTileM * 4) + i * (kTileM * 4)));
      }
      const int c_local = static_cast<int>(crank) * W;
      epilogue_chunk<32>(p.ep, acc, s_scale, s_shift, c_local, n_tile * BN + c_local, n < p.batch, n, oh, ow);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
    return;
  }

  if constexpr (SPLITK == 1) {
    // #1: every CTA of the cluster has its accumulator (its epilogue warps arrive only after the wait above), so the
    // leader's pipeline stages are idle and may receive the peers' partials.
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (crank != 0 && warp >= 2) {
      const int q = warp & 3;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t dst = mapa_u32(smem_u32(smem) + SplitKPart<BN>::OFF + (crank - 1) * SplitKPart<BN>::PEER_BYTES +
                                        (q * 32 + lane) * 4, 0);
      if constexpr (BN >= 32) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t acc[32];
          tmem_ld32(trow + c0, acc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) st_cluster_u32(dst + (c0 + i) * (kTileM * 4), acc[i]);
        }
      } else {
        uint32_t acc[16];
        tmem_ld16(trow, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) st_cluster_u32(dst + i * (kTileM * 4), acc[i]);
      }
    }
    tc_fence_before();
    cluster_sync_all();  // #2: the partials are visible in the leader's shared memory
    tc_fence_after();
  }

  if (warp >= 2 && (SPLITK == 0 || crank == 0)) {
    // ------------------------------ epilogue -----------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int dw = m & (p.tw - 1);
    const int dh = (m >> p.tw_log2) & (p.th - 1);
    const int dn = m >> (p.tw_log2 + p.th_log2);
    const int n = n0 + dn, oh = h0 + dh, ow = w0 + dw;
    const bool valid = n < p.batch;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // split-K leader: add the peers' partials of columns [c0, c0 + CH) of this thread's row, in rank order
    auto add_partials = [&](auto& acc, int c0) {
      if constexpr (SPLITK == 1) {
        const uint32_t src = smem_u32(smem) + SplitKPart<BN>::OFF + (q * 32 + lane) * 4 + c0 * (kTileM * 4);
        constexpr int CH = sizeof(acc) / sizeof(uint32_t);
        for (int r = 1; r < p.ksplit; ++r) {
#pragma unroll
          for (int i = 0; i < CH; ++i)
            acc[i] = __float_as_uint(__uint_as_float(acc[i]) +
                                     lds_f32(src + (r - 1) * SplitKPart<BN>::PEER_BYTES + i * (kTileM * 4)));
        }
      }
    };
    if constexpr (BN >= 32) {
      const bool bf16_out = (p.ep.flags & (VUB_CONV_F32_NHWC | VUB_CONV_F32_NCHW)) == 0;
      if (bf16_out) {
        // All MMAs (hence all TMA loads) of this CTA are complete: the pipeline stages are free and
        // serve as the per-warp staging tiles of the coalesced store path.
        constexpr int GW = BN < 128 ? BN : 128;
        const uint32_t st_raw = smem_u32(smem) + q * EpiStage<GW>::WARP_BYTES;
        const uint32_t st_elu = st_raw + EpiStage<GW>::TILE_BYTES;
        auto coord = [&](int row, int& rn, int& roh, int& row_w) {
          const int mm = q * 32 + row;
          rn = n0 + (mm >> (p.tw_log2 + p.th_log2));
          roh = h0 + ((mm >> p.tw_log2) & (p.th - 1));
          row_w = w0 + (mm & (p.tw - 1));
          return rn < p.batch;
        };
#pragma unroll 1
        for (int g0 = 0; g0 < BN; g0 += GW) {
#pragma unroll 1
          for (int c0 = g0; c0 < g0 + GW; c0 += 32) {
            uint32_t acc[32];
            tmem_ld32(trow + c0, acc);
            tmem_ld_wait();
            add_partials(acc, c0);
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaf(__uint_as_float(acc[i]), s_scale[c0 + i], s_shift[c0 + i]);
            if (p.ep.residual != nullptr && valid) {
              const long long pix = ((long long)n * p.ep.hout + oh) * p.ep.wout + ow;
              const uint4* r = reinterpret_cast<const uint4*>(p.ep.residual + pix * p.ep.cout + n_tile * BN + c0);
#pragma unroll
              for (int j = 0; j < 4; ++j) add_bf16x8(__ldg(r + j), &y[8 * j]);
            }
            if (p.ep.out_raw != nullptr) stage_put32<GW>(st_raw, lane, c0 - g0, y);
            if (p.ep.out_elu != nullptr) {
              const int mode = act_mode_of(p.ep.flags);
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] = act_apply(y[i], mode);
              stage_put32<GW>(st_elu, lane, c0 - g0, y);
            }
          }
          __syncwarp();
          if (p.ep.out_raw != nullptr) stage_flush<GW>(st_raw, p.ep.out_raw, p.ep, lane, n_tile * BN + g0, coord);
          if (p.ep.out_elu != nullptr) stage_flush<GW>(st_elu, p.ep.out_elu, p.ep, lane, n_tile * BN + g0, coord);
          __syncwarp();
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t acc[32];
          tmem_ld32(trow + c0, acc);
          tmem_ld_wait();
          add_partials(acc, c0);
          epilogue_chunk<32>(p.ep, acc, s_scale, s_shift, c0, n_tile * BN + c0, valid, n, oh, ow);
        }
      }
    } else {
      uint32_t acc[16];
      tmem_ld16(trow, acc);
      tmem_ld_wait();
      add_partials(acc, 0);
      epilogue_chunk<16>(p.ep, acc, s_scale, s_shift, 0, n_tile * BN, valid, n, oh, ow);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
template <int BN, int BKC, int SPLITK>
static int launch_conv_t(const ConvKParams& kp, dim3 grid, cudaStream_t stream) {
  constexpr int STAGE_BYTES = (kTileM + BN) * BKC * 2;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_fwd_kernel<BN, BKC, SPLITK>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "cudaFuncSetAttribute(conv_fwd<%d,%d>): %s", BN, BKC,
                cudaGetErrorString(attr_err));
  size_t smem = static_cast<size_t>(kp.stages) * STAGE_BYTES;
  constexpr int GW = BN < 128 ? BN : 128;
  constexpr size_t kStaging = BN >= 32 ? 4 * (size_t)EpiStage<(GW < 32 ? 32 : GW)>::WARP_BYTES : 0;
  if constexpr (kStaging > 0) {
    if (smem < kStaging) smem = kStaging;  // the epilogue reuses the (then idle) pipeline stages
  }
  if constexpr (SPLITK == 1) {  // the leader's stages also receive the peers' partial accumulators
    const size_t need = SplitKPart<BN>::OFF + (size_t)(kp.ksplit - 1) * SplitKPart<BN>::PEER_BYTES;
    if (smem < need) smem = need;
  }
  if constexpr (SPLITK == 2) {  // every CTA receives (ksplit - 1) slices of 128 x 32 fp32
    const size_t need = (size_t)(kp.ksplit - 1) * 32 * kTileM * 4;
    if (smem < need) smem = need;
  }
  smem += 1024;
  if (smem > 200 * 1024) return fail(VUB_E_SHAPE, "conv_fwd<%d,%d>: %zu bytes of shared memory", BN, BKC, smem);
  return launch_pdl(conv_fwd_kernel<BN, BKC, SPLITK>, grid, dim3(kConvThreads), smem, stream, kp,
                    SPLITK ? "conv_fwd_splitk" : "conv_fwd", SPLITK ? kp.ksplit : 1);
}

template <int BN, int BKC>
static int launch_conv(const ConvKParams& kp, dim3 grid, cudaStream_t stream) {
  if (kp.ksplit > 1 && kp.reduce_scatter) {
    if constexpr (BN >= 64 && BKC == 64) {
      if (BN == 32 * kp.ksplit) return launch_conv_t<BN, BKC, 2>(kp, grid, stream);
    }
    return fail(VUB_E_SHAPE, "conv_fwd: reduce-scatter split-K needs N tile = 32 x ranges (64 / 128 / 256)");
  }
  if (kp.ksplit > 1) {
    if constexpr (BN <= 128) return launch_conv_t<BN, BKC, 1>(kp, grid, stream);
    return fail(VUB_E_SHAPE, "conv_fwd: split-K needs an N tile <= 128");
  }
  return launch_conv_t<BN, BKC, 0>(kp, grid, stream);
}

bool conv_rows_applicable(const vub_conv_desc* d);            // conv_rows.cu
int conv_rows_launch(const vub_conv_desc* d, cudaStream_t stream);

static int default_stages(int bn, int bkc, int ksteps) {
  const int stage_bytes = (kTileM + bn) * bkc * 2;
  const int budget = (bn >= 128 ? 100 : 56) * 1024;
  int s = budget / stage_bytes;
  if (s < 2) s = 2;
  if (s > kMaxStages) s = kMaxStages;
  if (s > ksteps) s = ksteps;
  return s;
}

}  // namespace vub

using namespace vub;

extern "C" int vub_conv_fwd_elu_in_ok(int32_t nsrc, int32_t c_src, int32_t cout, int32_t ksize, int32_t stride,
                                      int32_t hin, int32_t win) {
  vub_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.nsrc = nsrc;
  d.c_src = c_src;
  d.cout = cout;
  d.ksize = ksize;
  d.stride = stride;
  d.hin = hin;
  d.win = win;
  d.flags = VUB_CONV_ELU_IN0;
  return conv_rows_applicable(&d) ? 1 : 0;
}

extern "C" int vub_conv_fwd_u8_out_ok(int32_t nsrc, int32_t c_src, int32_t cout, int32_t ksize, int32_t stride,
                                      int32_t hin, int32_t win) {
  vub_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.nsrc = nsrc;
  d.c_src = c_src;
  d.cout = cout;
  d.ksize = ksize;
  d.stride = stride;
  d.hin = hin;
  d.win = win;
  d.flags = VUB_CONV_U8_HWC;
  return (cout == 16 && conv_rows_applicable(&d)) ? 1 : 0;
}

extern "C" int vub_conv_fwd(const vub_conv_desc* d, void* stream_v) {
  clear_error();
  if (!d) return fail(VUB_E_NULL, "conv_fwd: desc is NULL");
  if (!d->x[0] || !d->w || !d->scale || !d->shift) return fail(VUB_E_NULL, "conv_fwd: x/w/scale/shift NULL");
  if (d->nsrc != 1 && d->nsrc != 2) return fail(VUB_E_SHAPE, "conv_fwd: nsrc must be 1 or 2");
  if (d->nsrc == 2 && !d->x[1]) return fail(VUB_E_NULL, "conv_fwd: x[1] NULL with nsrc == 2");
  // "linear layer": ksize 1 on a 1x1 map = nn.Linear / nn.LSTMCell rows (behavior net, models/pose_behavior_rnn.py,
  // lib/modules.py:236-257): any K that is a multiple of 64 up to 2048, any N that is a multiple of 128 up to 4096
  const bool linear = d->ksize == 1 && d->stride == 1 && d->hin == 1 && d->win == 1 && !(d->flags & VUB_CONV_D2S);
  if (d->c_src != 32 && d->c_src != 64 && d->c_src != 128 && d->c_src != 256 && d->c_src != 512 &&
      !(linear && d->c_src % 64 == 0 && d->c_src >= 64 && d->c_src <= 2048))
    return fail(VUB_E_SHAPE, "conv_fwd: c_src %d not in {32,64,128,256,512} (linear layers: multiples of 64 <= 2048)",
                d->c_src);
  // 256 / 512 channels per source: the 4-group autoregressive prior of VunetOrg (models/vunets.py:655-676), which
  // lives on the two smallest latent maps only -- served by the per-tap kernel (any number of channel blocks)
  const bool wide = d->c_src > 128 || d->cout > 512 || (d->flags & (VUB_CONV_LRELU | VUB_CONV_TANH)) != 0;
  if (wide && (d->stride != 1 || (d->flags & VUB_CONV_D2S)))
    return fail(VUB_E_SHAPE, "conv_fwd: c_src %d needs stride 1 and no D2S", d->c_src);
  if (d->ksize != 1 && d->ksize != 3) return fail(VUB_E_SHAPE, "conv_fwd: ksize must be 1 or 3");
  if (d->stride != 1 && d->stride != 2) return fail(VUB_E_SHAPE, "conv_fwd: stride must be 1 or 2");
  if (d->stride == 2 && (d->ksize != 3 || d->nsrc != 1))
    return fail(VUB_E_SHAPE, "conv_fwd: stride 2 needs ksize 3 and one source");
  if (d->batch <= 0 || !is_pow2(d->hin) || !is_pow2(d->win))
    return fail(VUB_E_SHAPE, "conv_fwd: batch %d hin %d win %d (extents must be powers of two)", d->batch,
                d->hin, d->win);
  const int hout = d->hin / d->stride, wout = d->win / d->stride;
  if (hout < 1 || wout < 1) return fail(VUB_E_SHAPE, "conv_fwd: empty output");
  const int cout = d->cout;
  if (!(cout == 16 || cout == 32 || cout == 64 || cout == 128 || cout == 256 || cout == 512 ||
        (linear && cout % 128 == 0 && cout > 0 && cout <= 4096)))
    return fail(VUB_E_SHAPE, "conv_fwd: cout %d not in {16,32,64,128,256,512} (linear layers: multiples of 128 <= 4096)",
                cout);
  if (d->cout_valid <= 0 || d->cout_valid > cout) return fail(VUB_E_SHAPE, "conv_fwd: cout_valid");
  const uint32_t f = d->flags;
  const bool f32 = (f & (VUB_CONV_F32_NHWC | VUB_CONV_F32_NCHW | VUB_CONV_U8_HWC)) != 0;  // "not bf16 NHWC"
  if (((f & VUB_CONV_F32_NHWC) != 0) + ((f & VUB_CONV_F32_NCHW) != 0) + ((f & VUB_CONV_U8_HWC) != 0) > 1)
    return fail(VUB_E_FLAGS, "conv_fwd: more than one of F32_NHWC / F32_NCHW / U8_HWC");
  if ((f & VUB_CONV_U8_HWC) && !(cout == 16 && conv_rows_applicable(d)))
    return fail(VUB_E_FLAGS, "conv_fwd: U8_HWC output is only implemented by the row-streaming kernel for cout 16 "
                             "(see vub_conv_fwd_u8_out_ok)");
  if (f32 && (!d->out_f32 || d->out_raw || d->out_elu || (f & VUB_CONV_D2S)))
    return fail(VUB_E_FLAGS, "conv_fwd: f32 output excludes bf16 outputs and D2S");
  if (!f32 && (d->out_f32 || (!d->out_raw && !d->out_elu)))
    return fail(VUB_E_FLAGS, "conv_fwd: need out_raw and/or out_elu (and no out_f32)");
  if (!(f & (VUB_CONV_F32_NCHW | VUB_CONV_U8_HWC)) && d->cout_valid != cout)
    return fail(VUB_E_FLAGS, "conv_fwd: cout_valid != cout only with F32_NCHW / U8_HWC");
  if ((f & VUB_CONV_D2S) && (d->residual || cout < 128))
    return fail(VUB_E_FLAGS, "conv_fwd: D2S excludes residual and needs cout >= 128");
  const void* ptrs[] = {d->x[0], d->x[1], d->w, d->residual, d->out_raw, d->out_elu, d->out_f32};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(VUB_E_ALIGN, "conv_fwd: pointer %p not 16B aligned", q);

  // Split-K over a thread-block cluster (per-tap kernel, partials reduced through distributed shared memory) for the
  // smallest maps.  Chosen from the LAYER geometry only -- never from the batch -- so that a frame's bits do not
  // depend on the batch it is rendered in (the partial sums change the fp32 summation order).
  int ksplit = 1;
  {
    // Measured on B200 at 64 frames (profiles/r02_splitk.log): on 4x4 maps (8 pixel tiles) three k-ranges per tile
    // beat the N split alone (128->128 3x3 0.010 -> 0.008 ms, [128|128]->128 0.016 -> 0.010 ms), and at 8 frames also on
    // 8x8 maps; on 8x8 maps at 64 frames (32 pixel tiles) the grid is already full and the clusters only add their
    // barriers (0.010 -> 0.017 ms).  The rule may not look at the batch, so it stops at 4x4.  VUB_SPLITK=0 disables
    // it, VUB_SPLITK=n >= 2 forces n ranges on every map up to 8x8 (experiments).
    const int ksteps_all = d->nsrc * d->ksize * d->ksize * (d->c_src / (d->c_src >= 64 ? 64 : 32));
    const int mode = env_int("VUB_SPLITK", 1);
    const bool eligible = ksteps_all >= 12 && !(f & (VUB_CONV_ELU_IN0 | VUB_CONV_ELU_IN1)) && d->stages >= 0;
    if (mode == 1 && eligible && hout * wout <= 16) ksplit = 3;
    // linear layers (batch rows as pixels: 2 pixel tiles at batch 256) with many output features already fill the chip
    // through their N tiles; measured (tools/linear_sweep.sh): 2048->2048 0.014 -> 0.013 ms, 1088->4096 0.021 -> 0.010 ms
    // without the split, 2048->512 0.013 -> 0.008 ms with it
    if (mode == 1 && linear && cout > 512) ksplit = 1;
    if (mode >= 2 && mode <= 8 && eligible && hout * wout <= 64) ksplit = mode;
  }
  if (ksplit == 1 && !wide && conv_rows_applicable(d)) return conv_rows_launch(d, static_cast<cudaStream_t>(stream_v));
  if (f & (VUB_CONV_ELU_IN0 | VUB_CONV_ELU_IN1))
    return fail(VUB_E_FLAGS, "conv_fwd: ELU-on-load is only implemented by the row-streaming kernel "
                             "(see vub_conv_fwd_elu_in_ok) — store an ELU copy for this layer");

  // tensor-core-bound 3x3 layers at 64x64 / 32x32: activation-slab kernel (A and weight reuse in smem)
  if (ksplit == 1 && !wide && !(f & VUB_CONV_RELU) && d->ksize == 3 && d->stride == 1 && d->stages >= 0 && !f32 && conv_slab_shape_ok(d->c_src, d->hin, d->win, cout, d->nsrc * 9 * d->c_src, ((long long)d->batch * d->hin * d->win + 127) / 128)) {
    SlabJob job;
    memset(&job, 0, sizeof(job));
    job.x[0] = d->x[0];
    job.x[1] = d->x[1];
    job.nsrc = d->nsrc;
    job.c_src = d->c_src;
    job.batch = d->batch;
    job.h = d->hin;
    job.w = d->win;
    job.w_packed = d->w;
    job.w_tiled = d->w_tiled;
    job.w_rows = cout;
    job.ktot = (long long)d->nsrc * 9 * d->c_src;
    job.n_total = cout;
    TileEpi& e = job.epi[0];
    e.scale = d->scale;
    e.shift = d->shift;
    e.addend = static_cast<const __nv_bfloat16*>(d->residual);
    e.out_raw = static_cast<__nv_bfloat16*>(d->out_raw);
    e.out_elu = static_cast<__nv_bfloat16*>(d->out_elu);
    const bool d2s_ = (f & VUB_CONV_D2S) != 0;
    e.out_h = d2s_ ? 2 * hout : hout;
    e.out_w = d2s_ ? 2 * wout : wout;
    e.out_c = d2s_ ? cout / 4 : cout;
    e.pix_stride = 1;
    e.n_valid = d->cout_valid;
    e.flags = f;
    return conv_slab_launch(job, static_cast<cudaStream_t>(stream_v));
  }

  // Measured at 64 frames (profiles/r01_tiles_vs_pertap.log): the persistent tile kernel wins for the
  // stride-2 and N >= 256 (up-sampling / latent) layers, the two-CTAs-per-SM per-tap kernel below for the
  // stride-1 N <= 128 layers.  VUB_FWD_TILES=1 / =0 forces one or the other.
  // On small maps (fewer CTAs than SMs) the persistent kernel's two or three 48 KB stages leave a lone CTA
  // latency bound (18 k-steps x ~2 us at 4x4); the per-tap kernel with its N split and deep pipeline is 3-5x
  // faster there.
  const char* force = getenv("VUB_FWD_TILES");
  const long long px_tiles = ((long long)d->batch * hout * wout + kTileM - 1) / kTileM;
  const bool small_grid = px_tiles * (cout > 256 ? cout / 256 : 1) < sm_count();
  // With the contiguous-tile epilogue (tile_epilogue_contig) the persistent kernel also wins the stride-1 N <= 128
  // layers (128->128 @64^2, B=64: 0.091 ms vs 0.136 ms per-tap), so everything with a full grid goes there.
  const bool want_tiles = ksplit == 1 && !wide && (force ? force[0] == '1' : !small_grid);
  if (want_tiles && d->stages >= 0) {
    // persistent tile kernel, k-step table = (source, tap, channel block)
    TileJob job;
    memset(&job, 0, sizeof(job));
    const int C = d->c_src;
    const int bkc_ = C >= 64 ? 64 : 32;
    const int cblocks = C / bkc_;
    const int taps = d->ksize * d->ksize;
    job.bkc = bkc_;
    job.w = d->w;
    job.w_tiled = bkc_ == 64 ? d->w_tiled : nullptr;
    job.w_rows = cout;
    job.ktot = (long long)d->nsrc * taps * C;
    job.n_total = cout;
    job.batch = d->batch;
    job.hout = hout;
    job.wout = wout;
    if (d->stride == 1) {
      job.n_src = d->nsrc;
      for (int sidx = 0; sidx < d->nsrc; ++sidx)
        job.src[sidx] = TileSrc{d->x[sidx], C, d->win, d->hin, (long long)C, (long long)d->win * C,
                                (long long)d->hin * d->win * C};
    } else {
      job.n_src = 4;
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw)
          job.src[ph * 2 + pw] =
              TileSrc{static_cast<const char*>(d->x[0]) + ((long long)ph * d->win + pw) * C * 2, C, d->win / 2,
                      d->hin / 2, 2LL * C, 2LL * d->win * C, (long long)d->hin * d->win * C};
    }
    int ks = 0;
    for (int sidx = 0; sidx < d->nsrc; ++sidx)
      for (int tap = 0; tap < taps; ++tap) {
        int dh = 0, dw = 0, mi = sidx;
        if (taps == 9) {
          const int r = tap / 3, q = tap % 3;
          if (d->stride == 1) {
            dh = r - 1;
            dw = q - 1;
          } else {  // input row 2*oh + r - 1: parity (r != 1), coordinate oh - (r == 0)
            mi = ((r != 1) ? 2 : 0) + ((q != 1) ? 1 : 0);
            dh = (r == 0) ? -1 : 0;
            dw = (q == 0) ? -1 : 0;
          }
        }
        for (int cb = 0; cb < cblocks; ++cb, ++ks)
          job.ksteps[ks] = KStep{(uint8_t)mi, (int8_t)dh, (int8_t)dw, (uint8_t)cb, (uint16_t)ks, 0};
      }
    job.n_ksteps = ks;
    TileEpi& e = job.epi[0];
    e.scale = d->scale;
    e.shift = d->shift;
    e.addend = static_cast<const __nv_bfloat16*>(d->residual);
    e.mask_elu = nullptr;
    e.out_raw = static_cast<__nv_bfloat16*>(d->out_raw);
    e.out_elu = static_cast<__nv_bfloat16*>(d->out_elu);
    e.out_f32 = static_cast<float*>(d->out_f32);
    const bool d2s = (f & VUB_CONV_D2S) != 0;
    e.out_h = d2s ? 2 * hout : hout;
    e.out_w = d2s ? 2 * wout : wout;
    e.out_c = d2s ? cout / 4 : cout;
    e.c_off = 0;
    e.pix_stride = 1;
    e.pix_off_h = e.pix_off_w = 0;
    e.n_valid = d->cout_valid;
    e.flags = f;
    return conv_tiles_launch(job, static_cast<cudaStream_t>(stream_v));
  }

  const int bkc = d->c_src >= 64 ? 64 : 32;
  int bn = cout > 256 ? 256 : cout;
  {
    // Small maps (<= 16x16): a handful of 128-pixel tiles leaves most SMs idle while each CTA streams the whole
    // K loop at the per-SM TMA rate.  Split N over more CTAs (each re-reads the A boxes, 4x smaller weight tiles)
    // until the grid covers about two thirds of the chip.
    const long long pix_tiles = ((long long)d->batch * hout * wout + kTileM - 1) / kTileM;
    while (bn > 32 && pix_tiles * (cout / bn) * ksplit < 96) bn >>= 1;
    // split-K: (ksplit - 1) partial tiles of 128 x bn fp32 must fit the leader's shared memory beside the staging
    // tiles (2, 3 -> bn <= 128, 4, 5, 6 -> 64, 7, 8 -> 32; a smaller N tile changes no bits)
    while (ksplit > 1 && bn > 16 && ((ksplit - 1) * bn > 320 || bn > 128)) bn >>= 1;
    while (cout % bn) bn >>= 1;  // linear layers: N is a multiple of 128, not necessarily of 256
  }
  // Linear layers with thousands of output features: split-K by reduce-scatter over the cluster (N tile = 32 x ranges,
  // every CTA owns 32 columns).  VUB_LINEAR_RS=0 disables it, 2 / 4 / 8 choose the range count.  Measured on the
  // behavior net's sampler (batch 256, one CUDA graph, profiles/r02_linear_rs.log): 3.22 ms without, 2.77 ms with two
  // ranges, 3.02 ms with four; eight-CTA clusters are slower than no split (2048->2048: 0.027 vs 0.015 ms) -- a cluster
  // needs all its SMs free in one GPC at once, and 7/8 of every accumulator crosses the cluster network.
  int reduce_scatter = 0;
  if (linear && cout > 512 && d->c_src >= 64 && d->stages >= 0) {
    const int want = env_int("VUB_LINEAR_RS", 2);
    const int ksteps_lin = d->nsrc * (d->c_src / 64);
    if ((want == 2 || want == 4 || want == 8) && ksteps_lin >= 2 * want && cout % (32 * want) == 0) {
      ksplit = want;
      bn = 32 * want;
      reduce_scatter = 1;
    }
  }

  ConvKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.ep.scale = d->scale;
  kp.ep.shift = d->shift;
  kp.ep.residual = static_cast<const __nv_bfloat16*>(d->residual);
  kp.ep.out_raw = static_cast<__nv_bfloat16*>(d->out_raw);
  kp.ep.out_elu = static_cast<__nv_bfloat16*>(d->out_elu);
  kp.ep.out_f32 = static_cast<float*>(d->out_f32);
  kp.ep.hout = hout;
  kp.ep.wout = wout;
  kp.ep.cout = cout;
  kp.ep.cout_valid = d->cout_valid;
  kp.ep.flags = f;
  kp.batch = d->batch;
  kp.nsrc = d->nsrc;
  kp.cblocks = d->c_src / bkc;
  kp.ntaps = d->ksize * d->ksize;
  kp.stride = d->stride;
  kp.tw = wout < kTileM ? wout : kTileM;
  kp.th = hout < kTileM / kp.tw ? hout : kTileM / kp.tw;
  kp.tn = kTileM / (kp.tw * kp.th);
  kp.tw_log2 = ilog2(kp.tw);
  kp.th_log2 = ilog2(kp.th);
  kp.tiles_w = wout / kp.tw;
  kp.tiles_h = hout / kp.th;
  kp.w_tiled = bkc == 64 ? static_cast<const uint8_t*>(d->w_tiled) : nullptr;
  kp.w_rows = cout;
  const int tiles_n = (d->batch + kp.tn - 1) / kp.tn;
  const int ksteps = kp.nsrc * kp.ntaps * kp.cblocks;
  kp.ksplit = ksplit;
  kp.reduce_scatter = reduce_scatter;
  const int ksteps_cta = (ksteps + ksplit - 1) / ksplit;  // the longest range of a split-K cluster
  kp.stages = d->stages > 0 ? d->stages : default_stages(bn, bkc, ksteps_cta);  // stages < 0: force this kernel
  if (kp.stages > ksteps_cta) kp.stages = ksteps_cta;
  // (not for linear layers: two of them -- the scale and the translation net of a coupling -- run side by side on two
  // streams and must be able to share an SM; VUB_LINEAR_DEEP=1 gives them the whole budget too)
  if (d->stages <= 0 && (!linear || env_int("VUB_LINEAR_DEEP", 0)) &&
      (long long)kp.tiles_w * kp.tiles_h * tiles_n * (cout / bn) * ksplit <= sm_count()) {
    // at most one CTA per SM: nothing to share the shared memory with, and with two stages only 40 KB are in
    // flight per L2/HBM round trip -> take the whole budget (the small-map layers are latency bound)
    const int deep = 190 * 1024 / ((kTileM + bn) * bkc * 2);
    kp.stages = deep < ksteps_cta ? deep : ksteps_cta;
  }
  if (kp.stages > kMaxStages) kp.stages = kMaxStages;
  if ((size_t)kp.stages * (kTileM + bn) * bkc * 2 + 1024 > 200 * 1024)
    return fail(VUB_E_SHAPE, "conv_fwd: %d stages do not fit in shared memory", kp.stages);

  // ---- tensor maps ----
  const uint64_t C = static_cast<uint64_t>(d->c_src);
  const uint64_t Hin = d->hin, Win = d->win, B = d->batch;
  const uint32_t boxA[4] = {(uint32_t)bkc, (uint32_t)kp.tw, (uint32_t)kp.th, (uint32_t)kp.tn};
  int rc;
  if (d->stride == 1) {
    const uint64_t dims[4] = {C, Win, Hin, B};
    const uint64_t strides[3] = {C * 2, Win * C * 2, Hin * Win * C * 2};
    for (int s = 0; s < d->nsrc; ++s) {
      rc = encode_tmap_bf16(&kp.tmA[s], d->x[s], 4, dims, strides, boxA, bkc * 2, false);
      if (rc) return rc;
    }
  } else {
    const uint64_t dims[4] = {C, Win / 2, Hin / 2, B};
    const uint64_t strides[3] = {2 * C * 2, 2 * Win * C * 2, Hin * Win * C * 2};
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const char* base = static_cast<const char*>(d->x[0]) + ((uint64_t)ph * Win + pw) * C * 2;
        rc = encode_tmap_bf16(&kp.tmA[ph * 2 + pw], base, 4, dims, strides, boxA, bkc * 2, false);
        if (rc) return rc;
      }
  }
  {
    const uint64_t ktot = (uint64_t)ksteps * bkc;
    const uint64_t dims[2] = {ktot, (uint64_t)cout};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t boxB[2] = {(uint32_t)bkc, (uint32_t)bn};
    rc = encode_tmap_bf16(&kp.tmB, d->w, 2, dims, strides, boxB, bkc * 2, true);
    if (rc) return rc;
  }

  dim3 grid(kp.tiles_w * kp.tiles_h * tiles_n, cout / bn, ksplit);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
#define VUB_DISPATCH(BN_, BKC_) \
  if (bn == BN_ && bkc == BKC_) return launch_conv<BN_, BKC_>(kp, grid, stream);
  VUB_DISPATCH(16, 32)
  VUB_DISPATCH(16, 64)
  VUB_DISPATCH(32, 32)
  VUB_DISPATCH(32, 64)
  VUB_DISPATCH(64, 32)
  VUB_DISPATCH(64, 64)
  VUB_DISPATCH(128, 32)
  VUB_DISPATCH(128, 64)
  VUB_DISPATCH(256, 32)
  VUB_DISPATCH(256, 64)
#undef VUB_DISPATCH
  return fail(VUB_E_SHAPE, "conv_fwd: no kernel for BN %d BKC %d", bn, bkc);
}
