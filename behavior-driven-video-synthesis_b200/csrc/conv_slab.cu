// Activation-slab implicit-GEMM kernel for the tensor-core-bound 3x3 stride-1 convs of the VUnet
// (>= 64 channels per source at 64x64 / 32x32: eu/du blocks 4-7, dd blocks 6-9, ups.3/ups.4) and for their
// data gradients.
//
// The per-tap kernels (conv_fwd.cu, conv_tiles.cu) pull a fresh 16 KB activation box AND a 16 KB weight
// tile through TMA for every 256 tensor-core cycles: 128 B/cycle/SM against the ~43 B/cycle/SM the
// L2 -> SM path delivers (profiles/r01_exp_rates_tma.log), so they top out near 1/3 of the MMA rate.
// Here one CTA owns a SUPER-TILE of 256 output pixels (R rows x W columns of one image = two 128-pixel
// M tiles) and, per (source, 64-channel block, column shift dw), loads ONE slab of (R+2) x W input pixels:
//   * the three row shifts (dh) of a tap column are three start addresses into the same slab (a row
//     shift is a multiple of W pixels = whole swizzle atoms),
//   * every weight tile is used by both M tiles.
// Bytes per MMA cycle drop from 128 to 62.5, i.e. the L2 -> SM path now sustains ~2/3 of the MMA rate.
// Accumulators: 2 M tiles x 128 columns, double-buffered in TMEM (512 columns); two epilogue warpgroups
// (one per M tile) drain buffer b while the tensor core fills buffer b^1.  The epilogue is the shared
// mode-driven one (tile_epilogue.cuh): forward (scale/shift, residual, raw/ELU, DepthToSpace) and data
// gradient (ELU' mask, two addends, per-source column split).
//
// Replaces NormConv2d.forward / its input gradient (lib/modules.py:140-145) for these shapes.
#include "conv_slab.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.h"
#include "conv_epilogue.cuh"
#include "ptx.cuh"
#include "tile_epilogue.cuh"

namespace vub {

constexpr int kSlabThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue of M tile 0, warps 6-9 of M tile 1
constexpr int kSlabABytes = 48 * 1024;  // one slab stage: (R+2) * W <= 384 pixels x 128 B
constexpr int kSlabBBytes = 16 * 1024;  // one weight tile: 128 rows x 64 channels
constexpr int kSlabMaxSA = 3;
constexpr int kSlabMaxSB = 6;
constexpr int kSlabBN = 128;
constexpr int kSlabGW = 32;   // channels staged per epilogue pass
constexpr int kSlabMaxN = 512;

struct SlabKParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  TileEpi ep[2];
  int n_split;      // N tiles [0, n_split) use ep[0]
  int stage_tiles;  // staging tiles per epilogue warp (2 or 3)
  int nsrc, cblocks;
  int W, H, batch, R, w_log2, th_log2;
  int tiles_h;      // H / R
  int n_ntiles, total_items, n_total;
  int flip;         // 0: forward (input pixel = output pixel + tap - 1); 1: data gradient (output - tap + 1)
  int sa, sb;       // ring depths
  const uint8_t* w_tiled;
  int w_rows;
  int contig;       // outputs written 1:1 (no DepthToSpace / scatter / fp32): lean linear-addressing epilogue
};

__global__ void __launch_bounds__(kSlabThreads, 1) conv_slab_kernel(const __grid_constant__ SlabKParams p) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t a_full[kSlabMaxSA], a_empty[kSlabMaxSA];
  __shared__ uint64_t b_full[kSlabMaxSB], b_empty[kSlabMaxSB];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_scale[kSlabMaxN];
  __shared__ float s_shift[kSlabMaxN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + p.sa * kSlabABytes;
  uint8_t* staging = b_ring + p.sb * kSlabBBytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.sa; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.sb; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);  // one arrival per epilogue warp (both M tiles)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.n_total; i += kSlabThreads) {
    s_scale[i] = p.ep[0].scale != nullptr ? p.ep[0].scale[i] : 1.0f;
    s_shift[i] = p.ep[0].shift != nullptr ? p.ep[0].shift[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int slab_bytes = (p.R + 2) * p.W * 128;
  const int tile_rows_bytes = (p.R / 2) * p.W * 128;  // = 128 pixels x 128 B = 16 KB
  auto item_coords = [&](int item, int& n_tile, int& n, int& h0) {
    n_tile = item % p.n_ntiles;
    const int st = item / p.n_ntiles;
    h0 = (st % p.tiles_h) * p.R;
    n = st / p.tiles_h;
  };

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int n_tile, n, h0;
      item_coords(item, n_tile, n, h0);
      for (int src = 0; src < p.nsrc; ++src)
        for (int cb = 0; cb < p.cblocks; ++cb)
          for (int dwi = 0; dwi < 3; ++dwi) {
            mbar_wait(&a_empty[sa], pha ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&a_full[sa], slab_bytes);
              tma_load_4d(a_ring + sa * kSlabABytes, &p.tmA[src], &a_full[sa], cb * 64, dwi - 1, h0 - 1, n);
            }
            __syncwarp();
            if (++sa == p.sa) {
              sa = 0;
              pha ^= 1;
            }
            for (int dhi = 0; dhi < 3; ++dhi) {
              const int tap = p.flip ? (2 - dhi) * 3 + (2 - dwi) : dhi * 3 + dwi;
              const int bblk = (src * 9 + tap) * p.cblocks + cb;
              mbar_wait(&b_empty[sb], phb ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&b_full[sb], kSlabBBytes);
                if (p.w_tiled != nullptr)  // one 1-D bulk copy instead of a 128-row tensor box
                  bulk_load_1d(b_ring + sb * kSlabBBytes, p.w_tiled + ((size_t)bblk * p.w_rows + n_tile * kSlabBN) * 128,
                               kSlabBBytes, &b_full[sb]);
                else
                  tma_load_2d(b_ring + sb * kSlabBBytes, &p.tmB, &b_full[sb], bblk * 64, n_tile * kSlabBN);
              }
              __syncwarp();
              if (++sb == p.sb) {
                sb = 0;
                phb ^= 1;
              }
            }
          }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(128, kSlabBN);
    const uint32_t desc_hi = umma_desc_hi_kmajor(128);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_ring));
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(b_ring));
    const int n_slabs = p.nsrc * p.cblocks * 3;
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    uint32_t li = 0;  // local item counter
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
      const uint32_t buf = li & 1;
      mbar_wait(&acc_empty[buf], ((li >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * (2 * kSlabBN);
      for (int s = 0; s < n_slabs; ++s) {
        mbar_wait(&a_full[sa], pha);
        const uint32_t a_lo = a_lo0 + sa * (kSlabABytes >> 4);
        for (int dhi = 0; dhi < 3; ++dhi) {
          mbar_wait(&b_full[sb], phb);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + sb * (kSlabBBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const uint32_t at = a_lo + ((t * tile_rows_bytes + dhi * p.W * 128) >> 4);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_bf16(d_tmem + t * kSlabBN, umma_desc_join(desc_hi, at + 2 * j), umma_desc_join(desc_hi, b_lo + 2 * j),
                          idesc, (s | dhi | j) != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[sb]);
            if (dhi == 2) {
              umma_commit(&a_empty[sa]);
              if (s == n_slabs - 1) umma_commit(&acc_full[buf]);
            }
          }
          __syncwarp();
          if (++sb == p.sb) {
            sb = 0;
            phb ^= 1;
          }
        }
        if (++sa == p.sa) {
          sa = 0;
          pha ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue: warpgroup g drains M tile g ----------------------
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const uint32_t t0 = smem_u32(staging) + (warp - 2) * p.stage_tiles * EpiStage<kSlabGW>::TILE_BYTES;
    const uint32_t t1 = t0 + EpiStage<kSlabGW>::TILE_BYTES;
    const uint32_t t2 = t1 + EpiStage<kSlabGW>::TILE_BYTES;
    const EpiGeom gm{p.W, p.R / 2, p.w_log2, p.th_log2, p.batch};
    uint32_t li = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++li) {
      int n_tile, n, h0;
      item_coords(item, n_tile, n, h0);
      const uint32_t buf = li & 1;
      const int ei = n_tile >= p.n_split ? 1 : 0;
      const TileEpi& e = p.ep[ei];
      const int ecol = (n_tile - (ei ? p.n_split : 0)) * kSlabBN;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (2 * kSlabBN) + g * kSlabBN;
      const bool affine = p.ep[0].scale != nullptr;  // data gradients run with the identity
      if (p.contig) {  // an M tile = R/2 full-width rows of one image = one contiguous run of the output
        const long long pix0 = ((long long)n * p.H + h0 + g * (p.R / 2)) * p.W + q * 32;
        tile_epilogue_contig<kSlabBN, kSlabGW>(e, affine ? s_scale + n_tile * kSlabBN : nullptr,
                                               affine ? s_shift + n_tile * kSlabBN : nullptr, trow, t0, t1, t2, lane, ecol,
                                               pix0, 32, &acc_empty[buf], &acc_full[buf], (li >> 1) & 1);
        continue;
      }
      tile_epilogue_staged<kSlabBN, kSlabGW>(e, gm, affine ? s_scale + n_tile * kSlabBN : nullptr,
                                             affine ? s_shift + n_tile * kSlabBN : nullptr, trow, t0, t1,
                                             t2, lane, q, ecol, n, h0 + g * (p.R / 2), 0, &acc_empty[buf], &acc_full[buf], (li >> 1) & 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
bool conv_slab_shape_ok(int c_src, int h, int w, int n_total, int k_total, long long pix_tiles) {
  // Opt-in (VUB_ENABLE_SLAB=1 takes every supported shape).  History, measured on B200: the kernel halves the
  // L2 -> SM traffic and beat the single-CTA tile kernel by 10 % on the concat layers (K >= 2304), where it was the
  // default; the CTA-pair tile kernel (conv_tiles.cu) now beats it there too ([128|128]->128 @32^2 B=64 forward:
  // 0.041 ms vs 0.045 ms) whenever the grid is full.  Choosing between the two by grid size would make a frame's bits
  // depend on the batch it is in (the slab kernel accumulates K in a different order than the tile kernels, which
  // agree bit for bit among themselves) -- tests/test_gpu_parity.py::test_batch_independence... -- so the tile
  // kernels take these layers at every batch size.
  (void)k_total;
  (void)pix_tiles;
  const char* env = getenv("VUB_ENABLE_SLAB");  // read per call: the native tests toggle it
  if (env == nullptr || env[0] != '1') return false;
  if (c_src % 64 || (w != 32 && w != 64)) return false;
  const int R = 256 / w;
  return h % R == 0 && n_total % kSlabBN == 0 && n_total <= kSlabMaxN;
}

int conv_slab_launch(const SlabJob& job, cudaStream_t stream) {
  if (!conv_slab_shape_ok(job.c_src, job.h, job.w, job.n_total, 1 << 30, 0))
    return fail(VUB_E_SHAPE, "conv_slab: unsupported shape c %d %dx%d N %d", job.c_src, job.h, job.w, job.n_total);
  if (job.n_split % kSlabBN) return fail(VUB_E_SHAPE, "conv_slab: split %d not a multiple of %d", job.n_split, kSlabBN);
  SlabKParams kp;
  memset(&kp, 0, sizeof(kp));
  const int n_epi = job.n_split > 0 && job.n_split < job.n_total ? 2 : 1;
  bool any_addend2 = false;
  for (int i = 0; i < n_epi; ++i) {
    if (job.epi[i].mask_elu && job.epi[i].out_elu)
      return fail(VUB_E_FLAGS, "conv_slab: mask and elu output share a staging tile");
    any_addend2 = any_addend2 || job.epi[i].addend2 != nullptr;
  }
  kp.ep[0] = job.epi[0];
  kp.ep[1] = job.epi[n_epi - 1];
  kp.n_split = n_epi == 2 ? job.n_split / kSlabBN : job.n_total / kSlabBN;
  kp.stage_tiles = any_addend2 ? 3 : 2;
  kp.nsrc = job.nsrc;
  kp.cblocks = job.c_src / 64;
  kp.W = job.w;
  kp.H = job.h;
  kp.batch = job.batch;
  kp.R = 256 / job.w;
  kp.w_log2 = ilog2(job.w);
  kp.th_log2 = ilog2(kp.R / 2);
  kp.tiles_h = job.h / kp.R;
  kp.n_ntiles = job.n_total / kSlabBN;
  kp.total_items = job.batch * kp.tiles_h * kp.n_ntiles;
  kp.n_total = job.n_total;
  kp.flip = job.flip;
  kp.contig = 1;
  for (int i = 0; i < n_epi; ++i) {
    const TileEpi& e = job.epi[i];
    if (e.flags != 0 || e.pix_stride != 1 || e.pix_off_h != 0 || e.pix_off_w != 0 || e.out_h != job.h || e.out_w != job.w ||
        e.out_f32 != nullptr)
      kp.contig = 0;
  }
  kp.w_tiled = static_cast<const uint8_t*>(job.w_tiled);
  kp.w_rows = job.w_rows;
  const size_t staging = (size_t)8 * kp.stage_tiles * EpiStage<kSlabGW>::TILE_BYTES;
  kp.sa = 2;
  kp.sb = static_cast<int>((222 * 1024 - 4608 - staging - (size_t)kp.sa * kSlabABytes) / kSlabBBytes);
  if (kp.sb > kSlabMaxSB) kp.sb = kSlabMaxSB;
  if (kp.sb < 3) return fail(VUB_E_SHAPE, "conv_slab: pipeline does not fit");
  const size_t smem = (size_t)kp.sa * kSlabABytes + (size_t)kp.sb * kSlabBBytes + staging + 1024;

  int rc;
  const uint64_t C = job.c_src, W = job.w, H = job.h, B = job.batch;
  const uint64_t dims[4] = {C, W, H, B};
  const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  const uint32_t boxA[4] = {64u, (uint32_t)job.w, (uint32_t)(kp.R + 2), 1u};
  for (int s = 0; s < job.nsrc; ++s) {
    rc = encode_tmap_bf16(&kp.tmA[s], job.x[s], 4, dims, strides, boxA, 128, false);
    if (rc) return rc;
  }
  {
    const uint64_t dimsb[2] = {(uint64_t)job.ktot, (uint64_t)job.n_total};
    const uint64_t stridesb[1] = {(uint64_t)job.ktot * 2};
    const uint32_t boxB[2] = {64u, (uint32_t)kSlabBN};
    rc = encode_tmap_bf16(&kp.tmB, job.w_packed, 2, dimsb, stridesb, boxB, 128, true);
    if (rc) return rc;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "cudaFuncSetAttribute(conv_slab): %s", cudaGetErrorString(attr_err));
  const int n_sm = sm_count();
  const int grid = kp.total_items < n_sm ? kp.total_items : n_sm;
  conv_slab_kernel<<<grid, kSlabThreads, smem, stream>>>(kp);
  return check_launch("conv_slab");
}

}  // namespace vub
