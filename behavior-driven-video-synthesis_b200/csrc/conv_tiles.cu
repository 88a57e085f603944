// Persistent, warp-specialised implicit-GEMM tile kernel on tcgen05 (sm_100a) driven by a k-step table.
//
//   D[128 pixels, BN] (TMEM fp32, double-buffered) += A[128, BKC] (smem) * B[BN, BKC]^T (smem)   per k-step
//
// One CTA per SM loops over (pixel tile, N tile) pairs.  Warp 0 streams k-steps through a shared-memory
// ring with TMA (A = a 4-D box of an NHWC activation tensor or of a strided parity view of one, shifted
// by the k-step's pixel offset, out-of-bounds zero-filled; B = a tile of the packed weight matrix),
// running ahead across tile boundaries; warp 1 issues tcgen05.mma into alternating TMEM buffers; two
// epilogue warpgroups alternate tiles, so tile i's epilogue overlaps tile i+1's main loop.
// The epilogue is table-free but mode-driven (TileEpi): fp32 scale/shift, optional ELU'-mask multiply
// (data gradients), optional addend (residual / upstream gradient), raw and/or ELU bf16 outputs staged
// through shared memory and written with coalesced 16-byte chunks, DepthToSpace or strided pixel
// scatter, or fp32 NHWC / NCHW outputs.
//
// conv_pair_kernel (below) is the same loop on a CTA pair: tcgen05.mma.cta_group::2 with M = 256 over two
// neighbouring pixel tiles, half of the weight tile staged per CTA; it takes the launches with BN >= 64, 64-channel
// k-steps, contiguous epilogues and an even number of pixel tiles.
//
// Users: vub_conv_fwd (every conv the row-streaming kernel does not take) and vub_conv_dgrad.
// Replaces NormConv2d.forward / its autograd backward w.r.t. the input (lib/modules.py:140-145) with the
// surrounding VunetRNB / Upsample / Downsample elementwise work fused in.
#include "conv_tiles.cuh"

#include <cstring>
#include <mutex>

#include "common.h"
#include "conv_epilogue.cuh"
#include "ptx.cuh"
#include "tile_epilogue.cuh"

namespace vub {

constexpr int kTilesThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 epilogue warpgroups
constexpr int kTilesMaxStages = 8;
constexpr int kTilesMaxN = 512;

struct TilesKParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  KStep ksteps[kTileMaxKSteps];
  TileEpi ep[2];
  int n_split;      // N tiles [0, n_split) use ep[0], the rest ep[1]
  int stage_tiles;  // staging tiles per epilogue warp: 2, or 3 when an addend2 is present
  int n_ksteps;
  int batch, hout, wout;
  int n_total;
  int tw, th, tn, tw_log2, th_log2;
  int tiles_w, tiles_h;
  int n_ntiles;
  int total_tiles;
  int stages;
  const uint8_t* w_tiled;
  int w_rows;
  long long total_pix;  // CONTIG kernels: pixels of the output tensor
  int d2s_w;            // CONTIG kernels: > 0 = DepthToSpace output, width of the conv's own output grid
};

// CONTIG: every tile is one contiguous run of the NHWC output (full-width tiles, stride 1, no DepthToSpace, bf16
// outputs) -> the lean linear-addressing epilogue.
template <int BN, int BKC, bool CONTIG>
__global__ void __launch_bounds__(kTilesThreads, 1)
    conv_tiles_kernel(const __grid_constant__ TilesKParams p) {
  constexpr int ROWB = BKC * 2;
  constexpr int A_BYTES = 128 * ROWB;
  constexpr int B_BYTES = BN * ROWB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  constexpr int GW = BN >= 64 ? 64 : 32;  // channels staged per epilogue pass
  constexpr bool STAGED = BN >= 32;
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024B alignment");

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full_bar[kTilesMaxStages];
  __shared__ uint64_t empty_bar[kTilesMaxStages];
  __shared__ uint64_t acc_full[2];
  __shared__ uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_scale[kTilesMaxN];
  __shared__ float s_shift[kTilesMaxN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;
  const int KS = p.n_ksteps;
  uint8_t* staging = smem + S * STAGE_BYTES;  // [8 epilogue warps][stage_tiles][32][GW*2+16]

  pdl_launch_dependents();  // programmatic dependent launch: see launch_pdl (common.h)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<TMEM_COLS>(&tmem_base_slot);
    tmem_relinquish();
  }
  pdl_wait();  // below: reads of what earlier kernels wrote
  for (int i = threadIdx.x; i < p.n_total; i += kTilesThreads) {
    s_scale[i] = p.ep[0].scale != nullptr ? p.ep[0].scale[i] : 1.0f;
    s_shift[i] = p.ep[0].shift != nullptr ? p.ep[0].shift[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  auto tile_coords = [&](int tile, int& n_tile, int& n0, int& h0, int& w0) {
    n_tile = tile % p.n_ntiles;
    int t = tile / p.n_ntiles;
    const int twi = t % p.tiles_w;
    t /= p.tiles_w;
    const int thi = t % p.tiles_h;
    const int tni = t / p.tiles_h;
    w0 = twi * p.tw;
    h0 = thi * p.th;
    n0 = tni * p.tn;
  };

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int n_tile, n0, h0, w0;
      tile_coords(tile, n_tile, n0, h0, w0);
      for (int ks = 0; ks < KS; ++ks) {
        const KStep k = p.ksteps[ks];
        mbar_wait(&empty_bar[st], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[st], STAGE_BYTES);
          uint8_t* a = smem + st * STAGE_BYTES;
          tma_load_4d(a, &p.tmA[k.map], &full_bar[st], k.cblk * BKC, w0 + k.dw, h0 + k.dh, n0);
          if (BKC == 64 && p.w_tiled != nullptr)  // one 1-D bulk copy instead of a BN-row tensor box
            bulk_load_1d(a + A_BYTES, p.w_tiled + ((size_t)k.bblk * p.w_rows + n_tile * BN) * 128, B_BYTES, &full_bar[st]);
          else
            tma_load_2d(a + A_BYTES, &p.tmB, &full_bar[st], k.bblk * BKC, n_tile * BN);
        }
        __syncwarp();
        if (++st == S) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    const uint32_t desc_hi = umma_desc_hi_kmajor(ROWB);
    const uint32_t smem_lo = umma_desc_lo(smem_u32(smem));
    int st = 0;
    uint32_t ph = 0;
    uint32_t lt = 0;  // local tile counter
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const uint32_t b = lt & 1;
      mbar_wait(&acc_empty[b], ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + b * BN;
      for (int ks = 0; ks < KS; ++ks) {
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t a_lo = smem_lo + st * (STAGE_BYTES >> 4);
        const uint32_t b_lo = a_lo + (A_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < BKC / 16; ++j)
            umma_bf16(d_tmem, umma_desc_join(desc_hi, a_lo + 2 * j), umma_desc_join(desc_hi, b_lo + 2 * j), idesc,
                      (ks | j) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[st]);
          if (ks == KS - 1) umma_commit(&acc_full[b]);
        }
        __syncwarp();
        if (++st == S) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue (two warpgroups, alternating tiles) ---------------
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int dw = m & (p.tw - 1);
    const int dh = (m >> p.tw_log2) & (p.th - 1);
    const int dn = m >> (p.tw_log2 + p.th_log2);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * BN;
    const uint32_t t0 = smem_u32(staging) + (warp - 2) * p.stage_tiles * EpiStage<GW>::TILE_BYTES;  // addend in / raw out
    const uint32_t t1 = t0 + EpiStage<GW>::TILE_BYTES;                              // mask in / elu out
    const uint32_t t2 = t1 + EpiStage<GW>::TILE_BYTES;                              // addend2 in (stage_tiles == 3)
    uint32_t cnt = 0;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      if ((lt & 1) != (uint32_t)g) continue;
      int n_tile, n0, h0, w0;
      tile_coords(tile, n_tile, n0, h0, w0);
      const int n = n0 + dn, oh = h0 + dh, ow = w0 + dw;
      const bool valid = n < p.batch;
      const int ei = n_tile >= p.n_split ? 1 : 0;
      const TileEpi& e = p.ep[ei];
      const int ecol = (n_tile - (ei ? p.n_split : 0)) * BN;  // first column of this tile as the epilogue numbers it
      const bool f32_out = (e.flags & (VUB_CONV_F32_NHWC | VUB_CONV_F32_NCHW)) != 0;
      mbar_wait(&acc_full[g], cnt & 1);
      tc_fence_after();
      if constexpr (STAGED && CONTIG) {
        const bool affine = p.ep[0].scale != nullptr;  // data gradients run with the identity
        const long long pix0 = ((long long)n0 * p.hout + h0) * p.wout + w0 + q * 32;
        const long long left = p.total_pix - pix0;
        tile_epilogue_contig<BN, GW>(e, affine ? s_scale + n_tile * BN : nullptr, affine ? s_shift + n_tile * BN : nullptr,
                                     trow, t0, t1, t2, lane, ecol, pix0, left > 32 ? 32 : static_cast<int>(left),
                                     &acc_empty[g], &acc_full[g], cnt & 1, p.d2s_w);
      } else if constexpr (STAGED) {
        const bool affine = p.ep[0].scale != nullptr;  // data gradients run with the identity
        tile_epilogue_staged<BN, GW>(e, EpiGeom{p.tw, p.th, p.tw_log2, p.th_log2, p.batch},
                                     affine ? s_scale + n_tile * BN : nullptr, affine ? s_shift + n_tile * BN : nullptr, trow, t0, t1, t2, lane, q, ecol, n0, h0, w0, &acc_empty[g]);
      } else {
        // BN == 16: fp32 NCHW output of out_conv-like layers
        uint32_t acc[16];
        tmem_ld16(trow, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[g]);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = ecol + i;
            const float y = fmaf(__uint_as_float(acc[i]), s_scale[n_tile * BN + i], s_shift[n_tile * BN + i]);
            if (e.flags & VUB_CONV_F32_NCHW) {
              if (col < e.n_valid) e.out_f32[tile_out_offset(e, col, n, oh, ow)] = y;
            } else if (e.flags & VUB_CONV_F32_NHWC) {
              e.out_f32[tile_out_offset(e, col, n, oh, ow)] = y;
            }
          }
        }
      }
      ++cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster work on two neighbouring pixel tiles and one N
// tile.  Each CTA stages its own 128 pixel rows of A but only HALF of the weight tile; the leader's MMA warp issues
// M = 256 instructions that read both CTAs' shared memory and write both CTAs' TMEM.  Per 128 x BN x 64 of MMA
// work a CTA's shared memory takes 16 KB + BN/2 * 128 B instead of 16 KB + BN * 128 B, and the tensor core
// fetches the B operand once per pair -- the shared-memory pipe is what bounds the 128-channel tile loop
// (DESIGN.md section 8).  Contiguous bf16 epilogues only (CONTIG of the single-CTA kernel); all building blocks
// were verified in tests/native/exp_cta2.cu first.
//
// Barriers: full_bar lives in the leader (its arrive.expect_tx counts both CTAs' bytes; both CTAs' TMA loads
// complete_tx on it -- the peer's bytes of a phase cannot arrive before the previous phase completed, because the
// peer refills a stage only after the leader's multicast commit released it); empty_bar / acc_full exist in both CTAs and are
// signalled by multicast commits; acc_empty lives in the leader (count 8: four epilogue warps of each CTA).
template <int BN, int BKC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTilesThreads, 1)
    conv_pair_kernel(const __grid_constant__ TilesKParams p) {
  constexpr int ROWB = BKC * 2;
  constexpr int A_BYTES = 128 * ROWB;
  constexpr int B_BYTES = (BN / 2) * ROWB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int GW = 64;
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024B alignment");

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full_bar[kTilesMaxStages];
  __shared__ uint64_t empty_bar[kTilesMaxStages];
  __shared__ uint64_t acc_full[2];
  __shared__ uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_scale[kTilesMaxN];
  __shared__ float s_shift[kTilesMaxN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int S = p.stages;
  const int KS = p.n_ksteps;
  uint8_t* staging = smem + S * STAGE_BYTES;

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<TMEM_COLS>(&tmem_base_slot);
  pdl_wait();
  for (int i = threadIdx.x; i < p.n_total; i += kTilesThreads) {
    s_scale[i] = p.ep[0].scale != nullptr ? p.ep[0].scale[i] : 1.0f;
    s_shift[i] = p.ep[0].shift != nullptr ? p.ep[0].shift[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers initialised and TMEM allocated before any remote arrive / MMA
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // work item -> this CTA's pixel tile (the pair's two tiles are neighbours) and the N tile
  auto item_coords = [&](int item, int& n_tile, int& n0, int& h0, int& w0) {
    n_tile = item % p.n_ntiles;
    int t = (item / p.n_ntiles) * 2 + static_cast<int>(rank);
    const int twi = t % p.tiles_w;
    t /= p.tiles_w;
    const int thi = t % p.tiles_h;
    const int tni = t / p.tiles_h;
    w0 = twi * p.tw;
    h0 = thi * p.th;
    n0 = tni * p.tn;
  };
  const int first = blockIdx.x >> 1, step = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------ TMA producer (one per CTA) ------------------------------
    int st = 0;
    uint32_t ph = 0;
    for (int item = first; item < p.total_tiles; item += step) {
      int n_tile, n0, h0, w0;
      item_coords(item, n_tile, n0, h0, w0);
      for (int ks = 0; ks < KS; ++ks) {
        const KStep k = p.ksteps[ks];
        mbar_wait(&empty_bar[st], ph ^ 1);
        if (elect_one()) {
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[st]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[st], 2 * STAGE_BYTES);  // both CTAs' bytes; the peer only sends data
          uint8_t* a = smem + st * STAGE_BYTES;
          tma_load_4d_pair(a, &p.tmA[k.map], leader_full, k.cblk * BKC, w0 + k.dw, h0 + k.dh, n0);
          tma_load_2d_pair(a + A_BYTES, &p.tmB, leader_full, k.bblk * BKC, n_tile * BN + static_cast<int>(rank) * (BN / 2));
        }
        __syncwarp();
        if (++st == S) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ----------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
      const uint32_t desc_hi = umma_desc_hi_kmajor(ROWB);
      const uint32_t smem_lo = umma_desc_lo(smem_u32(smem));
      int st = 0;
      uint32_t ph = 0;
      uint32_t lt = 0;
      for (int item = first; item < p.total_tiles; item += step, ++lt) {
        const uint32_t b = lt & 1;
        mbar_wait(&acc_empty[b], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * BN;
        for (int ks = 0; ks < KS; ++ks) {
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint32_t a_lo = smem_lo + st * (STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + (A_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < BKC / 16; ++j)
              umma_pair_bf16(d_tmem, umma_desc_join(desc_hi, a_lo + 2 * j), umma_desc_join(desc_hi, b_lo + 2 * j), idesc,
                             (ks | j) != 0 ? 1u : 0u);
            umma_pair_commit(&empty_bar[st]);
            if (ks == KS - 1) umma_pair_commit(&acc_full[b]);
          }
          __syncwarp();
          if (++st == S) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (two warpgroups, alternating items) -------------
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * BN;
    const uint32_t t0 = smem_u32(staging) + (warp - 2) * p.stage_tiles * EpiStage<GW>::TILE_BYTES;
    const uint32_t t1 = t0 + EpiStage<GW>::TILE_BYTES;
    const uint32_t t2 = t1 + EpiStage<GW>::TILE_BYTES;
    const uint32_t leader_empty = mapa_u32(smem_u32(&acc_empty[g]), 0);
    const bool affine = p.ep[0].scale != nullptr;
    uint32_t cnt = 0;
    uint32_t lt = 0;
    for (int item = first; item < p.total_tiles; item += step, ++lt) {
      if ((lt & 1) != (uint32_t)g) continue;
      int n_tile, n0, h0, w0;
      item_coords(item, n_tile, n0, h0, w0);
      const int ei = n_tile >= p.n_split ? 1 : 0;
      const TileEpi& e = p.ep[ei];
      const int ecol = (n_tile - (ei ? p.n_split : 0)) * BN;
      const long long pix0 = ((long long)n0 * p.hout + h0) * p.wout + w0 + q * 32;
      const long long left = p.total_pix - pix0;
      tile_epilogue_contig<BN, GW>(e, affine ? s_scale + n_tile * BN : nullptr, affine ? s_shift + n_tile * BN : nullptr, trow,
                                   t0, t1, t2, lane, ecol, pix0, left > 32 ? 32 : static_cast<int>(left), nullptr,
                                   &acc_full[g], cnt & 1, p.d2s_w, leader_empty);
      ++cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's shared memory and barriers stay alive until the leader is done with them
  if (warp == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}

template <int BN, int BKC>
static int launch_pair_t(const TilesKParams& kp, int grid, size_t smem, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_pair_kernel<BN, BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "cudaFuncSetAttribute(conv_pair<%d,%d>): %s", BN, BKC,
                cudaGetErrorString(attr_err));
  return launch_pdl(conv_pair_kernel<BN, BKC>, dim3(grid), dim3(kTilesThreads), smem, stream, kp, "conv_pair");
}

// ------------------------------------------------------------------------------------------------
template <int BN, int BKC, bool CONTIG>
static int launch_tiles_t(const TilesKParams& kp, int grid, size_t smem, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_tiles_kernel<BN, BKC, CONTIG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    220 * 1024);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "cudaFuncSetAttribute(conv_tiles<%d,%d>): %s", BN, BKC,
                cudaGetErrorString(attr_err));
  return launch_pdl(conv_tiles_kernel<BN, BKC, CONTIG>, dim3(grid), dim3(kTilesThreads), smem, stream, kp, "conv_tiles");
}

template <int BN, int BKC>
static int launch_tiles(const TilesKParams& kp, bool contig, int grid, size_t smem, cudaStream_t stream) {
  if constexpr (BN >= 32) {
    if (contig) return launch_tiles_t<BN, BKC, true>(kp, grid, smem, stream);
  }
  return launch_tiles_t<BN, BKC, false>(kp, grid, smem, stream);
}

int conv_tiles_launch(const TileJob& job, cudaStream_t stream) {
  const int bkc = job.bkc;
  if (bkc != 32 && bkc != 64) return fail(VUB_E_SHAPE, "conv_tiles: bkc %d", bkc);
  const int n_total = job.n_total;
  if (n_total > kTilesMaxN) return fail(VUB_E_SHAPE, "conv_tiles: N %d > %d", n_total, kTilesMaxN);
  int bn = job.bn > 0 ? job.bn : (n_total > 256 ? 256 : n_total);
  if (!(bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256) || n_total % bn)
    return fail(VUB_E_SHAPE, "conv_tiles: N %d not tileable by %d", n_total, bn);
  if (job.n_split % bn) return fail(VUB_E_SHAPE, "conv_tiles: split %d not a multiple of the N tile %d", job.n_split, bn);
  if (job.n_ksteps <= 0 || job.n_ksteps > kTileMaxKSteps) return fail(VUB_E_SHAPE, "conv_tiles: %d k-steps", job.n_ksteps);
  if (!is_pow2(job.hout) || !is_pow2(job.wout)) return fail(VUB_E_SHAPE, "conv_tiles: tile grid %dx%d", job.hout, job.wout);
  const int n_epi = job.n_split > 0 && job.n_split < n_total ? 2 : 1;
  bool any_addend2 = false;
  for (int i = 0; i < n_epi; ++i) {
    if (job.epi[i].mask_elu && job.epi[i].out_elu)
      return fail(VUB_E_FLAGS, "conv_tiles: mask and elu output share a staging tile");
    any_addend2 = any_addend2 || job.epi[i].addend2 != nullptr;
  }

  TilesKParams kp;
  memset(&kp, 0, sizeof(kp));
  memcpy(kp.ksteps, job.ksteps, sizeof(KStep) * job.n_ksteps);
  kp.ep[0] = job.epi[0];
  kp.ep[1] = job.epi[n_epi - 1];
  kp.n_split = n_epi == 2 ? job.n_split / bn : n_total / bn;
  kp.stage_tiles = any_addend2 ? 3 : 2;
  kp.n_ksteps = job.n_ksteps;
  kp.w_tiled = bkc == 64 ? static_cast<const uint8_t*>(job.w_tiled) : nullptr;
  kp.w_rows = job.w_rows;
  kp.batch = job.batch;
  kp.hout = job.hout;
  kp.wout = job.wout;
  kp.n_total = n_total;
  kp.tw = job.wout < 128 ? job.wout : 128;
  kp.th = job.hout < 128 / kp.tw ? job.hout : 128 / kp.tw;
  kp.tn = 128 / (kp.tw * kp.th);
  kp.tw_log2 = ilog2(kp.tw);
  kp.th_log2 = ilog2(kp.th);
  kp.tiles_w = job.wout / kp.tw;
  kp.tiles_h = job.hout / kp.th;
  const int tiles_n = (job.batch + kp.tn - 1) / kp.tn;
  kp.n_ntiles = n_total / bn;
  kp.total_tiles = kp.tiles_w * kp.tiles_h * tiles_n * kp.n_ntiles;

  // contiguous tiles: full-width pixel tiles written 1:1 (no shuffle, no scatter, bf16)
  bool contig = kp.tw == job.wout && !getenv("VUB_TILES_NO_CONTIG");
  bool d2s_all = true;
  const int gw_epi = bn >= 64 ? 64 : 32;
  for (int i = 0; i < n_epi; ++i) {
    const TileEpi& e = job.epi[i];
    const bool plain = (e.flags & ~VUB_CONV_RELU) == 0 && e.out_h == job.hout && e.out_w == job.wout;
    // DepthToSpace: a warp's 32 pixels inside one image row, a channel group inside one quadrant
    const bool d2s = (e.flags & ~VUB_CONV_RELU) == VUB_CONV_D2S && e.out_h == 2 * job.hout && e.out_w == 2 * job.wout && kp.tw >= 32 &&
                     e.out_c % gw_epi == 0 && e.c_off == 0 && n_epi == 1;
    d2s_all = d2s_all && d2s;
    contig = contig && (plain || d2s) && e.pix_stride == 1 && e.pix_off_h == 0 && e.pix_off_w == 0 && e.out_f32 == nullptr;
  }
  kp.d2s_w = (contig && d2s_all) ? job.wout : 0;
  kp.total_pix = (long long)job.batch * job.hout * job.wout;

  // CTA pairs (conv_pair_kernel): contiguous epilogues, an even number of pixel tiles, enough items for every pair
  const int pix_tiles = kp.tiles_w * kp.tiles_h * tiles_n;
  const int pair_env = getenv("VUB_PAIR") ? atoi(getenv("VUB_PAIR")) : 1;  // VUB_PAIR=0: single-CTA tiles only
  const bool pair = pair_env != 0 && contig && bkc == 64 && (bn == 64 || bn == 128 || bn == 256) && pix_tiles % 2 == 0;
  if (pair) kp.total_tiles = (pix_tiles / 2) * kp.n_ntiles;  // work items of a pair

  const int gw = bn >= 64 ? 64 : 32;
  const size_t staging = bn >= 32 ? (size_t)8 * kp.stage_tiles * 32 * (gw * 2) : 0;
  const size_t stage_bytes = (size_t)(128 + (pair ? bn / 2 : bn)) * bkc * 2;
  int stages = static_cast<int>((212 * 1024 - staging) / stage_bytes);
  if (stages > kTilesMaxStages) stages = kTilesMaxStages;
  if (stages < 2) return fail(VUB_E_SHAPE, "conv_tiles: pipeline does not fit");
  kp.stages = stages;
  const size_t smem = stages * stage_bytes + staging + 1024;

  const uint32_t boxA[4] = {(uint32_t)bkc, (uint32_t)kp.tw, (uint32_t)kp.th, (uint32_t)kp.tn};
  int rc;
  for (int s = 0; s < job.n_src; ++s) {
    const TileSrc& t = job.src[s];
    const uint64_t dims[4] = {(uint64_t)t.channels, (uint64_t)t.w, (uint64_t)t.h, (uint64_t)job.batch};
    const uint64_t strides[3] = {(uint64_t)t.stride_w * 2, (uint64_t)t.stride_h * 2, (uint64_t)t.stride_n * 2};
    rc = encode_tmap_bf16(&kp.tmA[s], t.base, 4, dims, strides, boxA, bkc * 2, false);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)job.ktot, (uint64_t)n_total};
    const uint64_t strides[1] = {(uint64_t)job.ktot * 2};
    const uint32_t boxB[2] = {(uint32_t)bkc, (uint32_t)(pair ? bn / 2 : bn)};
    rc = encode_tmap_bf16(&kp.tmB, job.w, 2, dims, strides, boxB, bkc * 2, true);
    if (rc) return rc;
  }
  const int n_sm = sm_count();
  const int grid = kp.total_tiles < n_sm ? kp.total_tiles : n_sm;
  if (pair) {
    const int pairs = kp.total_tiles < n_sm / 2 ? kp.total_tiles : n_sm / 2;
    kp.w_tiled = nullptr;
    if (bn == 64) return launch_pair_t<64, 64>(kp, 2 * pairs, smem, stream);
    if (bn == 128) return launch_pair_t<128, 64>(kp, 2 * pairs, smem, stream);
    return launch_pair_t<256, 64>(kp, 2 * pairs, smem, stream);
  }
#define VUB_TILES_DISPATCH(BN_, BKC_) \
  if (bn == BN_ && bkc == BKC_) return launch_tiles<BN_, BKC_>(kp, contig, grid, smem, stream);
  VUB_TILES_DISPATCH(16, 32)
  VUB_TILES_DISPATCH(16, 64)
  VUB_TILES_DISPATCH(32, 32)
  VUB_TILES_DISPATCH(32, 64)
  VUB_TILES_DISPATCH(64, 32)
  VUB_TILES_DISPATCH(64, 64)
  VUB_TILES_DISPATCH(128, 32)
  VUB_TILES_DISPATCH(128, 64)
  VUB_TILES_DISPATCH(256, 32)
  VUB_TILES_DISPATCH(256, 64)
#undef VUB_TILES_DISPATCH
  return fail(VUB_E_SHAPE, "conv_tiles: no kernel for BN %d BKC %d", bn, bkc);
}

}  // namespace vub
