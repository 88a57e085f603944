// Weight correlation of the weight-normalised convs (vub_conv_wgrad) on tcgen05 (sm_100a).
//
//   gt[split][co][(src, tap, ci)] = sum_{pix in split} dy[pix, co] * x_src[pix + off(tap), ci]        (fp32)
//   dy_sum[split][co]             = sum_{pix in split} dy[pix, co]
//
// The contraction index is the PIXEL, the slow index of both NHWC operands, so both UMMA operands are
// MN-major: a TMA box [channels, pixels] written with 128B/64B swizzle IS the canonical MN-major
// operand layout (LBO = stride between 64-/32-channel boxes, SBO = 8 pixel rows; measured:
// profiles/r01_exp_mnmajor.log).  One k-step = 128 pixels (64 on tiny maps):
//   A = dy box(es)  [128 output channels x 128 pixels]   (M = 128; layers with fewer output channels
//                                                         replicate their box with a block stride of 0)
//   B = x boxes of the CTA's tap group.  For 3x3 stride-1 convs at width >= 16 ONE box per (tap row,
//       channel block) is loaded with a one-pixel halo on both sides, and the three column taps are three
//       N blocks of the SAME box whose block stride (LBO) is one pixel row: a pixel shift is a start-address
//       offset in the MN-major layout (profiles/r01_exp_mnmajor.log), so x is read 3x instead of 9x.
//       Otherwise (1x1, stride 2, tiny maps) every tap is its own shifted box = one block of D columns.
//   D[co, (tap, ci)] (TMEM, fp32) accumulates over ALL pixel tiles the CTA owns; an extra N = 16 block of
//   constant ones yields dy_sum from the same A operand.
// Work split: grid = (tap group) x (128-channel M tile) x (pixel range = split); every CTA stores its partial
// correlation into the split's slice of the workspace with coalesced plain stores; vub_wn_bwd sums the splits
// (deterministic, no atomics: fp32 reductions to global memory were 3x the cost of the MMA main loop).
// Upsample convs read dy through the four parity maps of the shuffled gradient; Downsample convs read x
// through its four parity maps (same tap -> map/shift table as the forward kernel).
//
// Replaces autograd's convolution_backward (weight gradient) of lib/modules.py:143.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.h"
#include "ptx.cuh"

namespace vub {

constexpr int kWgThreads = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kWgMaxKP = 128;     // pixels per k-step: 128 (TMA costs >= ~384 cycles per box: 64-pixel boxes
                                  // made the stream box-count bound), 64 for tiny maps
constexpr int kWgMaxStages = 16;
constexpr int kWgMaxGroups = 6;
constexpr int kWgMaxBoxes = 9;    // B boxes per tap group
constexpr int kWgOnesBytes = 2048;

struct WgBBox {
  int16_t map;       // B tensor map (source or parity view)
  int8_t dh, dw;     // pixel shift of the box origin relative to the dy tile (halo boxes: dw = -1)
  int16_t chan;      // first channel of the box
  int32_t gt_row;    // gt row of the box's first column (halo boxes: of column tap 0; + c_src per column tap)
};
struct WgABox {
  int16_t map;
  int16_t chan;
};

struct WgKParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB[4];
  WgABox abox[4 * 4];                          // [m_tile][box]
  WgBBox bbox[kWgMaxGroups][kWgMaxBoxes];
  int nb[kWgMaxGroups];                        // boxes per group
  float* gt;             // partial correlations [n_split][cout_k][gt_pitch]
  float* dy_sum;         // partial sums [n_split][cout_k]
  int gt_pitch;
  long long split_stride;  // floats between the gt slices of consecutive splits
  int sum_stride;          // floats between the dy_sum slices of consecutive splits
  int ca, na;            // channels per A box, boxes per M tile (na * ca <= 128)
  int m_rows;            // valid rows of an M tile (min(cout_k, 128))
  int cb;                // channels per B box
  int halo;              // 1: B boxes carry a one-pixel halo and serve the 3 column taps (ndw = 3)
  int c_src;             // gt row stride between column taps of a halo box
  int b_box_pitch;       // bytes between consecutive B boxes of a stage (1024-aligned)
  int b_box_bytes;       // bytes TMA writes per B box
  int n_groups, n_mtiles, n_split;
  int kp_pix;            // pixels per k-step (64 or 128)
  int tw, th, tn;        // pixel tile (tw * th * tn = kp_pix)
  int tiles_w, tiles_h;
  int total_tiles;
  int stages;
};

// Shared-memory descriptor of an MN-major operand made of TMA boxes [channels x pixels] (swizzle = row
// bytes): LBO = byte stride between consecutive channel blocks (boxes), SBO = 8 pixel rows.
__device__ __forceinline__ uint32_t umma_desc_hi_mn(uint32_t row_bytes) {
  const uint32_t layout = (row_bytes == 128) ? 2u : 4u;
  return ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint32_t umma_desc_lo_mn(uint32_t smem_addr, uint32_t block_stride) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((block_stride >> 4) << 16);
}

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgKParams p) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full_bar[kWgMaxStages];
  __shared__ uint64_t empty_bar[kWgMaxStages];
  __shared__ uint64_t acc_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;

  // work assignment: blockIdx.x = (split * n_mtiles + mt) * n_groups + grp
  int bid = blockIdx.x;
  const int grp = bid % p.n_groups;
  bid /= p.n_groups;
  const int mt = bid % p.n_mtiles;
  const int split = bid / p.n_mtiles;
  const int t_begin = static_cast<int>((long long)p.total_tiles * split / p.n_split);
  const int t_end = static_cast<int>((long long)p.total_tiles * (split + 1) / p.n_split);
  const int nb = p.nb[grp];
  const bool with_sum = grp == 0;

  const int a_rowb = p.ca * 2, b_rowb = p.cb * 2;
  const int a_box = p.kp_pix * a_rowb, b_box = p.b_box_pitch;
  const int a_bytes = p.na * a_box;
  const int stage_bytes = a_bytes + nb * p.b_box_bytes;  // bytes TMA delivers per stage
  const int stage_pitch = (a_bytes + nb * b_box + 1023) & ~1023;
  const int ndw = p.halo ? 3 : 1;
  uint8_t* ones = smem;                 // 2 KB of bf16 1.0
  uint8_t* ring = smem + kWgOnesBytes;  // [S][stage_pitch]

  pdl_launch_dependents();  // programmatic dependent launch: see launch_pdl (common.h)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kWgOnesBytes / 4; i += kWgThreads)
    reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;  // two bf16 1.0
  // make the generic-proxy writes of `ones` visible to the tensor-core (async) proxy
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // below: reads of what earlier kernels wrote (dy, x) and writes to the shared partial workspace
  const uint32_t tmem_base = tmem_base_slot;
  const int n_cols = nb * ndw * p.cb;  // D columns of this CTA (the ones block follows at column n_cols)

  auto tile_origin = [&](int tile, int& n0, int& h0, int& w0) {
    const int twi = tile % p.tiles_w;
    int t = tile / p.tiles_w;
    const int thi = t % p.tiles_h;
    const int tni = t / p.tiles_h;
    w0 = twi * p.tw;
    h0 = thi * p.th;
    n0 = tni * p.tn;
  };

  if (t_begin < t_end) {
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      int st = 0;
      uint32_t ph = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        int n0, h0, w0;
        tile_origin(tile, n0, h0, w0);
        mbar_wait(&empty_bar[st], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[st], stage_bytes);
          uint8_t* dst = ring + st * stage_pitch;
          for (int i = 0; i < p.na; ++i) {
            const WgABox ab = p.abox[mt * 4 + i];
            tma_load_4d(dst + i * a_box, &p.tmA[ab.map], &full_bar[st], ab.chan, w0, h0, n0);
          }
          for (int j = 0; j < nb; ++j) {
            const WgBBox bb = p.bbox[grp][j];
            tma_load_4d(dst + a_bytes + j * b_box, &p.tmB[bb.map], &full_bar[st], bb.chan, w0 + bb.dw, h0 + bb.dh, n0);
          }
        }
        __syncwarp();
        if (++st == S) {
          st = 0;
          ph ^= 1;
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer --------------------------------
      // Everything that does not depend on the stage is hoisted out of the loop: the issuing thread is a single
      // lane, and descriptor arithmetic with divisions / parameter loads inside the loop made it the bottleneck
      // (ncu source view: ~340 instructions per K slice, profiles/r01_wgrad_issue_bound.txt).
      const uint32_t a_hi = umma_desc_hi_mn(a_rowb), b_hi = umma_desc_hi_mn(b_rowb);
      const uint32_t a_blk = p.na * p.ca >= 128 ? a_box : 0;  // fewer than 128 channels: replicate the box(es)
      const uint32_t a_lo_c = (a_blk >> 4) << 16;
      const uint64_t d_ones = umma_desc_join(umma_desc_hi_mn(128), umma_desc_lo_mn(smem_u32(ones), 0));
      constexpr uint32_t kMajorMN = (1u << 15) | (1u << 16);
      constexpr uint32_t idesc_ones = umma_idesc_bf16(128, 16) | kMajorMN;
      // per-CTA MMA list of one K slice: at most 3 MMAs (halo: one per box; otherwise chunks of <= 256 columns)
      uint32_t m_rel[3] = {0u, 0u, 0u}, m_col[3] = {0u, 0u, 0u}, m_idesc[3] = {0u, 0u, 0u}, m_lo_c[3] = {0u, 0u, 0u};
      int n_m = 0;
      if (p.halo) {
        for (int bx = 0; bx < nb && bx < 3; ++bx, ++n_m) {
          m_rel[bx] = bx * b_box;
          m_col[bx] = bx * 3 * p.cb;
          m_idesc[bx] = umma_idesc_bf16(128, 3 * p.cb) | kMajorMN;
          m_lo_c[bx] = (static_cast<uint32_t>(b_rowb) >> 4) << 16;  // block stride = one pixel row
        }
      } else {
        const int max_box = 256 / p.cb;
        for (int box0 = 0; box0 < nb && n_m < 3; ++n_m) {
          const int nbox = nb - box0 < max_box ? nb - box0 : max_box;
          m_rel[n_m] = box0 * b_box;
          m_col[n_m] = box0 * p.cb;
          m_idesc[n_m] = umma_idesc_bf16(128, nbox * p.cb) | kMajorMN;
          m_lo_c[n_m] = (static_cast<uint32_t>(b_box) >> 4) << 16;
          box0 += nbox;
        }
      }
      constexpr int kMaxSlices = kWgMaxKP / 16;
      const int n_slices = p.kp_pix / 16;
      uint32_t a_off[kMaxSlices], b_off[kMaxSlices];  // byte offsets of the K slices (16 pixels each) inside a box
#pragma unroll
      for (int j = 0; j < kMaxSlices; ++j) {
        const int px = j * 16;
        a_off[j] = px * a_rowb;
        b_off[j] = p.halo ? ((px / p.tw) * (p.tw + 2) + (px % p.tw)) * b_rowb : px * b_rowb;
      }
      const uint32_t ring_addr = smem_u32(ring);
      int st = 0;
      uint32_t ph = 0;
      uint32_t acc = 0u;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t a_addr = ring_addr + st * stage_pitch;
        const uint32_t b_addr = a_addr + a_bytes;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < kMaxSlices; ++j) {
            if (j >= n_slices) break;
            const uint64_t da = umma_desc_join(a_hi, (((a_addr + a_off[j]) & 0x3FFFFu) >> 4) | a_lo_c);
#pragma unroll
            for (int i = 0; i < 3; ++i)
              if (i < n_m)
                umma_bf16(tmem_base + m_col[i], da,
                          umma_desc_join(b_hi, (((b_addr + m_rel[i] + b_off[j]) & 0x3FFFFu) >> 4) | m_lo_c[i]), m_idesc[i],
                          acc);
            if (with_sum) umma_bf16(tmem_base + n_cols, da, d_ones, idesc_ones, acc);
            acc = 1u;
          }
          umma_commit(&empty_bar[st]);
          if (tile == t_end - 1) umma_commit(&acc_full);
        }
        __syncwarp();
        acc = 1u;
        if (++st == S) {
          st = 0;
          ph ^= 1;
        }
      }
    } else {
      // ------------------------------ epilogue: TMEM -> this split's partial correlation ----------------------
      // Plain coalesced stores into the split's private slice of the workspace (vub_wn_bwd sums the splits):
      // fp32 reductions to global memory cost ~4 cycles per lane and dominated the kernel (128 x 384 values per
      // CTA), and the result is now bit-reproducible.  Each 32x32 accumulator block is transposed through shared
      // memory (the pipeline ring is idle by now) so that a warp writes 128 contiguous bytes of one gt row.
      const int q = warp & 3;
      const int row0 = q * 32;  // first D row (= output channel within the M tile) of this warp
      mbar_wait(&acc_full, 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(row0) << 16);
      float* tp = reinterpret_cast<float*>(ring) + q * (32 * 33);
      float* gt_split = p.gt + (size_t)split * p.split_stride;
      for (int j = 0; j < nb; ++j) {
        for (int dwi = 0; dwi < ndw; ++dwi) {
          const int gt_col = p.bbox[grp][j].gt_row + dwi * p.c_src;
          for (int c0 = 0; c0 < p.cb; c0 += 32) {
            uint32_t acc[32];
            tmem_ld32(trow + (j * ndw + dwi) * p.cb + c0, acc);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) tp[lane * 33 + i] = __uint_as_float(acc[i]);
            __syncwarp();
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const int row = row0 + r;
              if (row < p.m_rows) gt_split[(size_t)(mt * 128 + row) * p.gt_pitch + gt_col + c0 + lane] = tp[r * 33 + lane];
            }
            __syncwarp();
          }
        }
      }
      if (with_sum) {
        uint32_t acc[16];
        tmem_ld16(trow + n_cols, acc);
        tmem_ld_wait();
        const int row = row0 + lane;
        if (row < p.m_rows) p.dy_sum[(size_t)split * p.sum_stride + mt * 128 + row] = __uint_as_float(acc[0]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace vub

using namespace vub;

// A launch may cover a 128-channel SLICE of wider sources (VunetOrg's 256 / 512-channel prior convs): `d` then
// describes the slice as a one-source conv, and this says where the slice sits in the real tensors.
struct WgSlice {
  int s_base;   // index of the slice's source among the conv's sources (gt rows are ordered (source, tap, channel))
  int c_total;  // channels per pixel of the x tensors = channels per (source, tap) block of gt
  int c_off;    // first channel of the slice
};

// Fills everything of the launch parameters that depends on shapes only (no pointers, no tensor maps) and
// returns the best number of pixel splits in *best_split.
static int wgrad_plan_kp(const vub_wgrad_desc* d, WgKParams& kp, int* best_split, int* max_nb_out, int kp_pix,
                         const WgSlice& sl) {
  if (d->nsrc != 1 && d->nsrc != 2) return fail(VUB_E_SHAPE, "conv_wgrad: nsrc must be 1 or 2");
  const int C = d->c_src;
  if (C != 32 && C != 64 && C != 128) return fail(VUB_E_SHAPE, "conv_wgrad: c_src %d not in {32,64,128}", C);
  if (d->ksize != 1 && d->ksize != 3) return fail(VUB_E_SHAPE, "conv_wgrad: ksize must be 1 or 3");
  if (d->stride != 1 && d->stride != 2) return fail(VUB_E_SHAPE, "conv_wgrad: stride must be 1 or 2");
  const bool d2s = (d->flags & VUB_CONV_D2S) != 0;
  if (d->stride == 2 && (d->ksize != 3 || d->nsrc != 1 || d2s))
    return fail(VUB_E_SHAPE, "conv_wgrad: stride 2 needs ksize 3, one source, no D2S");
  if (d->batch <= 0 || !is_pow2(d->hin) || !is_pow2(d->win) || d->hin < d->stride || d->win < d->stride)
    return fail(VUB_E_SHAPE, "conv_wgrad: batch %d hin %d win %d", d->batch, d->hin, d->win);
  const int cout_k = d->cout_k;
  if (cout_k < 32 || cout_k > 512 || cout_k % 32 || (cout_k > 32 && cout_k % 64) || (cout_k > 64 && cout_k % 128))
    return fail(VUB_E_SHAPE, "conv_wgrad: cout_k %d", cout_k);
  const int cmap = d2s ? cout_k / 4 : cout_k;
  if (cmap % 32) return fail(VUB_E_SHAPE, "conv_wgrad: %d channels per dy map", cmap);
  const int taps = d->ksize * d->ksize;
  const int ho = d->hin / d->stride, wo = d->win / d->stride;  // dy grid == tile grid (conv output pixels)

  memset(&kp, 0, sizeof(kp));
  // ---- A side: dy ----
  kp.ca = cmap >= 64 ? 64 : 32;
  const int m_rows = cout_k < 128 ? cout_k : 128;
  kp.m_rows = m_rows;
  kp.na = m_rows / kp.ca;
  kp.n_mtiles = (cout_k + 127) / 128;
  for (int mt = 0; mt < kp.n_mtiles; ++mt)
    for (int i = 0; i < kp.na; ++i) {
      const int co0 = mt * 128 + i * kp.ca;
      kp.abox[mt * 4 + i] = WgABox{(int16_t)(d2s ? co0 / cmap : 0), (int16_t)(d2s ? co0 % cmap : co0)};
    }
  // ---- pixel tiles of 64 ----
  kp.kp_pix = kp_pix;
  const int KP = kp.kp_pix;
  kp.tw = wo < KP ? wo : KP;
  kp.th = ho < KP / kp.tw ? ho : KP / kp.tw;
  kp.tn = KP / (kp.tw * kp.th);
  kp.tiles_w = wo / kp.tw;
  kp.tiles_h = ho / kp.th;
  const int tiles_n = (d->batch + kp.tn - 1) / kp.tn;
  kp.total_tiles = kp.tiles_w * kp.tiles_h * tiles_n;

  // ---- B side: x, grouped by (source, tap row) or (source) ----
  kp.cb = C >= 64 ? 64 : 32;
  kp.c_src = sl.c_total;  // gt row stride between the column taps of a halo box (== C unless a channel slice)
  const int cblocks = C / kp.cb;
  const bool merge_rows = taps == 9 && 9 * C <= 448;  // all 9 taps of a source fit one CTA's TMEM
  static const bool no_halo = getenv("VUB_WGRAD_NO_HALO") != nullptr && getenv("VUB_WGRAD_NO_HALO")[0] == '1';
  kp.halo = (taps == 9 && d->stride == 1 && kp.tw >= 16 && !no_halo) ? 1 : 0;
  kp.b_box_bytes = (kp.tw + 2 * kp.halo) * kp.th * kp.tn * kp.cb * 2;
  kp.b_box_pitch = (kp.b_box_bytes + 1023) & ~1023;
  int g = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    const int n_rowgroups = (taps == 9 && !merge_rows) ? 3 : 1;
    for (int rg = 0; rg < n_rowgroups; ++rg, ++g) {
      if (g >= kWgMaxGroups) return fail(VUB_E_SHAPE, "conv_wgrad: too many tap groups");
      int j = 0;
      if (kp.halo) {
        // one halo box per (tap row, channel block); its 3 N blocks are the column taps q = 0, 1, 2
        for (int r = 0; r < 3; ++r) {
          if (n_rowgroups == 3 && r != rg) continue;
          for (int cbk = 0; cbk < cblocks; ++cbk, ++j)
            kp.bbox[g][j] = WgBBox{(int16_t)s, (int8_t)(r - 1), (int8_t)-1, (int16_t)(cbk * kp.cb),
                                   ((sl.s_base + s) * taps + r * 3) * sl.c_total + sl.c_off + cbk * kp.cb};
        }
        kp.nb[g] = j;
        continue;
      }
      for (int tap = 0; tap < taps; ++tap) {
        const int r = tap / d->ksize, q = tap % d->ksize;
        if (n_rowgroups == 3 && r != rg) continue;
        int dh = 0, dw = 0, mi = s;
        if (taps == 9) {
          if (d->stride == 1) {
            dh = r - 1;
            dw = q - 1;
          } else {  // input pixel 2*o + r - 1: parity map (r != 1, q != 1), coordinate o - (r == 0)
            mi = ((r != 1) ? 2 : 0) + ((q != 1) ? 1 : 0);
            dh = (r == 0) ? -1 : 0;
            dw = (q == 0) ? -1 : 0;
          }
        }
        for (int cbk = 0; cbk < cblocks; ++cbk, ++j) {
          if (j >= kWgMaxBoxes) return fail(VUB_E_SHAPE, "conv_wgrad: too many boxes in a tap group");
          kp.bbox[g][j] = WgBBox{(int16_t)mi, (int8_t)dh, (int8_t)dw, (int16_t)(cbk * kp.cb),
                                 ((sl.s_base + s) * taps + tap) * sl.c_total + sl.c_off + cbk * kp.cb};
        }
      }
      kp.nb[g] = j;
    }
  }
  kp.n_groups = g;
  int max_nb = 0;
  for (int i = 0; i < g; ++i) max_nb = kp.nb[i] > max_nb ? kp.nb[i] : max_nb;
  const int cols_per_box = (kp.halo ? 3 : 1) * kp.cb;
  if (max_nb * cols_per_box + 16 > 512)
    return fail(VUB_E_SHAPE, "conv_wgrad: %d accumulator columns", max_nb * cols_per_box);

  const int n_sm = sm_count();
  int n_split = n_sm / (kp.n_groups * kp.n_mtiles);
  // Every split writes (and vub_wn_bwd reads back) a full [cout_k x K] fp32 partial: on the small maps the
  // partials of 49 splits were several times the layer's own input bytes.  Keep at least kMinSteps pixel
  // k-steps per split so that a CTA's MMA time is comparable to the time it needs to store its partial.
  static const int min_steps = getenv("VUB_WGRAD_MIN_STEPS") ? atoi(getenv("VUB_WGRAD_MIN_STEPS")) : 4;
  if (min_steps > 1 && n_split > kp.total_tiles / min_steps) n_split = kp.total_tiles / min_steps;
  if (n_split < 1) n_split = 1;
  if (n_split > kp.total_tiles) n_split = kp.total_tiles;
  *best_split = n_split;
  *max_nb_out = max_nb;
  return 0;
}

static size_t wgrad_stage_pitch(const WgKParams& kp, int max_nb) {
  const size_t stage_bytes = (size_t)kp.na * kp.kp_pix * kp.ca * 2 + (size_t)max_nb * kp.b_box_pitch;
  return (stage_bytes + 1023) & ~(size_t)1023;
}

// 128-pixel k-steps when the map is large enough and at least 3 pipeline stages fit, else 64-pixel k-steps.
static int wgrad_plan(const vub_wgrad_desc* d, WgKParams& kp, int* best_split, int* max_nb_out, const WgSlice& sl) {
  static const bool kp64 = getenv("VUB_WGRAD_KP64") != nullptr && getenv("VUB_WGRAD_KP64")[0] == '1';
  const long long pixels = (long long)d->batch * (d->hin / (d->stride > 0 ? d->stride : 1)) *
                           (d->win / (d->stride > 0 ? d->stride : 1));
  if (pixels >= 4 * kWgMaxKP && !kp64) {
    const int rc = wgrad_plan_kp(d, kp, best_split, max_nb_out, kWgMaxKP, sl);
    if (rc) return rc;
    if ((200 * 1024 - kWgOnesBytes) / wgrad_stage_pitch(kp, *max_nb_out) >= 3) return 0;
  }
  return wgrad_plan_kp(d, kp, best_split, max_nb_out, 64, sl);
}

// The slice of source `s` starting at channel `c0` of a wide conv, described as a one-source 128-channel conv.
static vub_wgrad_desc wgrad_slice_desc(const vub_wgrad_desc* d, int s, int c0) {
  vub_wgrad_desc sd = *d;
  sd.nsrc = 1;
  sd.c_src = 128;
  sd.x[0] = d->x[s] ? static_cast<const char*>(d->x[s]) + (size_t)c0 * 2 : nullptr;
  sd.x[1] = nullptr;
  return sd;
}

static int wgrad_check_wide(const vub_wgrad_desc* d) {
  if (d->nsrc != 1 && d->nsrc != 2) return fail(VUB_E_SHAPE, "conv_wgrad: nsrc must be 1 or 2");
  if (d->stride != 1 || (d->flags & VUB_CONV_D2S))
    return fail(VUB_E_SHAPE, "conv_wgrad: c_src %d needs stride 1 and no D2S", d->c_src);
  return 0;
}

extern "C" int32_t vub_conv_wgrad_splits(const vub_wgrad_desc* d) {
  clear_error();
  if (!d) return fail(VUB_E_NULL, "conv_wgrad_splits: desc is NULL");
  WgKParams kp;
  int best = 0, max_nb = 0;
  if (d->c_src == 256 || d->c_src == 512) {  // wide sources: the split count of one 128-channel slice launch
    if (int rcw = wgrad_check_wide(d)) return rcw;
    const vub_wgrad_desc sd = wgrad_slice_desc(d, 0, 0);
    const int rc = wgrad_plan(&sd, kp, &best, &max_nb, WgSlice{0, d->c_src, 0});
    return rc ? rc : best;
  }
  const int rc = wgrad_plan(d, kp, &best, &max_nb, WgSlice{0, d->c_src, 0});
  return rc ? rc : best;
}

static int wgrad_launch(const vub_wgrad_desc* d, const WgSlice& sl, void* stream_v);

extern "C" int vub_conv_wgrad(const vub_wgrad_desc* d, void* stream_v) {
  clear_error();
  if (!d) return fail(VUB_E_NULL, "conv_wgrad: desc is NULL");
  if (!d->dy || !d->x[0] || !d->gt || !d->dy_sum) return fail(VUB_E_NULL, "conv_wgrad: dy/x/gt/dy_sum NULL");
  if (d->nsrc == 2 && !d->x[1]) return fail(VUB_E_NULL, "conv_wgrad: x[1] NULL with nsrc == 2");
  if (d->c_src == 256 || d->c_src == 512) {
    // VunetOrg's 4-group prior (models/vunets.py:655-676): one launch per 128-channel slice of every source; each
    // writes its own (source, tap, channel) rows of gt (all of them write the same dy_sum values)
    if (int rcw = wgrad_check_wide(d)) return rcw;
    if (d->gt_pitch < d->nsrc * d->ksize * d->ksize * d->c_src)
      return fail(VUB_E_SHAPE, "conv_wgrad: gt_pitch %d < K", d->gt_pitch);
    for (int s = 0; s < d->nsrc; ++s)
      for (int c0 = 0; c0 < d->c_src; c0 += 128) {
        const vub_wgrad_desc sd = wgrad_slice_desc(d, s, c0);
        const int rc = wgrad_launch(&sd, WgSlice{s, d->c_src, c0}, stream_v);
        if (rc) return rc;
      }
    return 0;
  }
  return wgrad_launch(d, WgSlice{0, d->c_src, 0}, stream_v);
}

static int wgrad_launch(const vub_wgrad_desc* d, const WgSlice& sl, void* stream_v) {
  WgKParams kp;
  int best = 0, max_nb = 0;
  int rc = wgrad_plan(d, kp, &best, &max_nb, sl);
  if (rc) return rc;
  const int C = d->c_src, cout_k = d->cout_k;
  const bool d2s = (d->flags & VUB_CONV_D2S) != 0;
  const int cmap = d2s ? cout_k / 4 : cout_k;
  const int ho = d->hin / d->stride, wo = d->win / d->stride;
  const int ktot = d->nsrc * d->ksize * d->ksize * sl.c_total;
  if (d->gt_pitch < ktot) return fail(VUB_E_SHAPE, "conv_wgrad: gt_pitch %d < K = %d", d->gt_pitch, ktot);
  if (d->n_split < 1 || d->n_split > kp.total_tiles || d->n_split > 1024)
    return fail(VUB_E_SHAPE, "conv_wgrad: n_split %d not in [1, %d] (vub_conv_wgrad_splits suggests %d)", d->n_split,
                kp.total_tiles < 1024 ? kp.total_tiles : 1024, best);
  if (d->n_split > 1 && (d->split_stride < (int64_t)cout_k * d->gt_pitch || d->sum_stride < cout_k))
    return fail(VUB_E_SHAPE, "conv_wgrad: split strides too small");
  const void* ptrs[] = {d->dy, d->x[0], d->x[1]};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(VUB_E_ALIGN, "conv_wgrad: pointer %p not 16B aligned", q);
  kp.gt = d->gt;
  kp.dy_sum = d->dy_sum;
  kp.gt_pitch = d->gt_pitch;
  kp.split_stride = d->split_stride;
  kp.sum_stride = d->sum_stride;
  kp.n_split = d->n_split;

  const size_t stage_pitch = wgrad_stage_pitch(kp, max_nb);
  int stages = static_cast<int>((200 * 1024 - kWgOnesBytes) / stage_pitch);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  if (stages < 2) return fail(VUB_E_SHAPE, "conv_wgrad: pipeline does not fit");
  kp.stages = stages;
  const size_t smem = kWgOnesBytes + stages * stage_pitch + 1024;

  // ---- tensor maps ----
  {
    const uint32_t boxA[4] = {(uint32_t)kp.ca, (uint32_t)kp.tw, (uint32_t)kp.th, (uint32_t)kp.tn};
    if (!d2s) {
      const uint64_t dims[4] = {(uint64_t)cout_k, (uint64_t)wo, (uint64_t)ho, (uint64_t)d->batch};
      const uint64_t str[3] = {(uint64_t)cout_k * 2, (uint64_t)wo * cout_k * 2, (uint64_t)ho * wo * cout_k * 2};
      rc = encode_tmap_bf16(&kp.tmA[0], d->dy, 4, dims, str, boxA, kp.ca * 2, false);
      if (rc) return rc;
    } else {
      const uint64_t W = wo, H = ho;
      const uint64_t dims[4] = {(uint64_t)cmap, W, H, (uint64_t)d->batch};
      const uint64_t str[3] = {2ull * cmap * 2, 4ull * W * cmap * 2, 4ull * H * W * cmap * 2};
      for (int b1 = 0; b1 < 2; ++b1)
        for (int b2 = 0; b2 < 2; ++b2) {
          const char* base = static_cast<const char*>(d->dy) + ((uint64_t)b1 * 2 * W + b2) * cmap * 2;
          rc = encode_tmap_bf16(&kp.tmA[b1 * 2 + b2], base, 4, dims, str, boxA, kp.ca * 2, false);
          if (rc) return rc;
        }
    }
    const uint32_t boxB[4] = {(uint32_t)kp.cb, (uint32_t)(kp.tw + 2 * kp.halo), (uint32_t)kp.th, (uint32_t)kp.tn};
    const uint64_t Cc = sl.c_total, Hin = d->hin, Win = d->win;  // channel pitch of a pixel (== C unless a slice)
    if (d->stride == 1) {
      const uint64_t dims[4] = {(uint64_t)C, Win, Hin, (uint64_t)d->batch};
      const uint64_t str[3] = {Cc * 2, Win * Cc * 2, Hin * Win * Cc * 2};
      for (int s = 0; s < d->nsrc; ++s) {
        rc = encode_tmap_bf16(&kp.tmB[s], d->x[s], 4, dims, str, boxB, kp.cb * 2, false);
        if (rc) return rc;
      }
    } else {
      const uint64_t dims[4] = {Cc, Win / 2, Hin / 2, (uint64_t)d->batch};
      const uint64_t str[3] = {2 * Cc * 2, 2 * Win * Cc * 2, Hin * Win * Cc * 2};
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const char* base = static_cast<const char*>(d->x[0]) + ((uint64_t)ph * Win + pw) * Cc * 2;
          rc = encode_tmap_bf16(&kp.tmB[ph * 2 + pw], base, 4, dims, str, boxB, kp.cb * 2, false);
          if (rc) return rc;
        }
    }
  }

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "cudaFuncSetAttribute(conv_wgrad): %s", cudaGetErrorString(attr_err));
  const int grid = kp.n_groups * kp.n_mtiles * kp.n_split;
  return launch_pdl(conv_wgrad_kernel, dim3(grid), dim3(kWgThreads), smem, static_cast<cudaStream_t>(stream_v), kp,
                    "conv_wgrad");
}
