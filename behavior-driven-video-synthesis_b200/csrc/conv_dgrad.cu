// Data gradient of the weight-normalised convs (vub_conv_dgrad): host-side construction of k-step
// tables for the persistent tile kernel (conv_tiles.cu).
//
//   dx_s[pix, ci] = sum_{tap, co} dy[pix - off(tap), co] * (scale[co] * v_hat[co, s, tap, ci])
//
// is the same implicit GEMM as the forward conv with the roles of the channel axes swapped: A = boxes of
// dy shifted by the NEGATED tap offset, B = the transposed, pre-scaled weight matrix `wd` that
// vub_wn_prep emits, N = the input channels of all sources.  The two halves of a VunetRNB's concat get
// their own epilogue (different ELU' mask, addends and output tensor) through the kernel's column split.
//   * Upsample (conv + DepthToSpace): dy is the gradient of the SHUFFLED tensor; the four parity
//     sub-grids of it are four strided tensor maps, one per output-channel quadrant.
//   * Downsample (stride 2): each input-pixel parity class (ph, pw) sees a fixed subset of the 3x3 taps
//     (1, 2, 2 or 4 of them); one launch per class, scattering to every second input pixel.
//
// Replaces autograd's convolution_backward (input gradient) of lib/modules.py:143 plus the ELU / cat /
// add backward kernels of lib/modules.py:225-233.
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "conv_slab.cuh"
#include "conv_tiles.cuh"

namespace vub {
bool conv_rows_dgrad(const vub_dgrad_desc* d, cudaStream_t stream, int* rc);  // conv_rows.cu
}
using namespace vub;

extern "C" int vub_conv_dgrad(const vub_dgrad_desc* d, void* stream_v) {
  clear_error();
  if (!d) return fail(VUB_E_NULL, "conv_dgrad: desc is NULL");
  if (!d->dy || !d->wd || !d->dx[0]) return fail(VUB_E_NULL, "conv_dgrad: dy/wd/dx[0] NULL");
  if (d->nsrc != 1 && d->nsrc != 2) return fail(VUB_E_SHAPE, "conv_dgrad: nsrc must be 1 or 2");
  if (d->nsrc == 2 && !d->dx[1]) return fail(VUB_E_NULL, "conv_dgrad: dx[1] NULL with nsrc == 2");
  const int C = d->c_src;
  if (C != 32 && C != 64 && C != 128 && C != 256 && C != 512)
    return fail(VUB_E_SHAPE, "conv_dgrad: c_src %d not in {32,64,128,256,512}", C);
  if (d->ksize != 1 && d->ksize != 3) return fail(VUB_E_SHAPE, "conv_dgrad: ksize must be 1 or 3");
  if (d->stride != 1 && d->stride != 2) return fail(VUB_E_SHAPE, "conv_dgrad: stride must be 1 or 2");
  const bool d2s = (d->flags & VUB_CONV_D2S) != 0;
  if (d->stride == 2 && (d->ksize != 3 || d->nsrc != 1 || d2s))
    return fail(VUB_E_SHAPE, "conv_dgrad: stride 2 needs ksize 3, one source, no D2S");
  if (d->batch <= 0 || !is_pow2(d->hin) || !is_pow2(d->win) || d->hin < d->stride || d->win < d->stride)
    return fail(VUB_E_SHAPE, "conv_dgrad: batch %d hin %d win %d", d->batch, d->hin, d->win);
  const int cout_k = d->cout_k;
  if (cout_k < 32 || cout_k > 512 || cout_k % 32) return fail(VUB_E_SHAPE, "conv_dgrad: cout_k %d", cout_k);
  const int cmap = d2s ? cout_k / 4 : cout_k;  // channels of one dy tensor map
  if (cmap % 32 || (cmap > 32 && cmap % 64)) return fail(VUB_E_SHAPE, "conv_dgrad: %d channels per dy map", cmap);
  const void* ptrs[] = {d->dy, d->wd, d->dx[0], d->dx[1], d->mask_elu[0], d->mask_elu[1], d->addend[0],
                        d->addend[1], d->addend2[0], d->addend2[1]};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(VUB_E_ALIGN, "conv_dgrad: pointer %p not 16B aligned", q);

  const int taps = d->ksize * d->ksize;
  const int bkc = cmap >= 64 ? 64 : 32;
  const int cblocks = cmap / bkc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

  if (C > 128) {
    // 256 / 512 channels per source (VunetOrg's 4-group prior, models/vunets.py:655-676, on the two smallest latent
    // maps): one launch of the tile kernel per 128-channel slice of every source -- the slice's rows of `wd`, and an
    // epilogue that addresses its channels inside the C-channel pixel (c_off).
    if (d->stride != 1 || d2s) return fail(VUB_E_SHAPE, "conv_dgrad: c_src %d needs stride 1 and no D2S", C);
    if (taps * cblocks > kTileMaxKSteps) return fail(VUB_E_SHAPE, "conv_dgrad: too many k-steps");
    const long long H = d->hin, W = d->win;
    const long long ktot = (long long)taps * cout_k;
    for (int s = 0; s < d->nsrc; ++s)
      for (int c0 = 0; c0 < C; c0 += 128) {
        TileJob job;
        memset(&job, 0, sizeof(job));
        job.bkc = bkc;
        job.w = static_cast<const char*>(d->wd) + ((long long)(s * C + c0) * ktot) * 2;
        job.w_rows = 128;
        job.ktot = ktot;
        job.n_total = 128;
        job.bn = 128;
        const long long pix_tiles = ((long long)d->batch * H * W + 127) / 128;
        while (job.bn > 32 && pix_tiles * (128 / job.bn) < 96) job.bn >>= 1;
        job.batch = d->batch;
        job.hout = d->hin;
        job.wout = d->win;
        job.n_src = 1;
        job.src[0] = TileSrc{d->dy, cout_k, d->win, d->hin, (long long)cout_k, W * cout_k, H * W * cout_k};
        TileEpi& e = job.epi[0];
        e.addend = static_cast<const __nv_bfloat16*>(d->addend[s]);
        e.addend2 = static_cast<const __nv_bfloat16*>(d->addend2[s]);
        e.mask_elu = static_cast<const __nv_bfloat16*>(d->mask_elu[s]);
        e.out_raw = static_cast<__nv_bfloat16*>(d->dx[s]);
        e.out_h = d->hin;
        e.out_w = d->win;
        e.out_c = C;
        e.c_off = c0;
        e.pix_stride = 1;
        e.n_valid = 128;
        e.flags = d->flags & VUB_CONV_RELU;
        int ks = 0;
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / d->ksize, q = tap % d->ksize;
          const int dh = taps == 9 ? 1 - r : 0, dw = taps == 9 ? 1 - q : 0;
          for (int cb = 0; cb < cblocks; ++cb, ++ks)
            job.ksteps[ks] = KStep{0, (int8_t)dh, (int8_t)dw, (uint8_t)cb, (uint16_t)((tap * cout_k + cb * bkc) / bkc), 0};
        }
        job.n_ksteps = ks;
        const int rc = conv_tiles_launch(job, stream);
        if (rc) return rc;
      }
    return 0;
  }

  TileJob job;
  memset(&job, 0, sizeof(job));
  job.bkc = bkc;
  job.w = d->wd;
  job.w_tiled = d->wd_tiled;  // rows = input channels of all sources, K = (tap, output channel)
  job.w_rows = d->nsrc * C;
  job.ktot = (long long)taps * cout_k;
  job.n_total = d->nsrc * C;
  job.bn = C;
  {
    // small maps: a handful of 128-pixel tiles cannot fill 148 SMs and each CTA would walk all k-steps behind a
    // 2-3 stage pipeline.  Split N (more CTAs, 20 KB stages -> 8-deep pipeline), as the forward per-tap kernel does.
    const long long pix_tiles = ((long long)d->batch * (d->hin / d->stride) * (d->win / d->stride) + 127) / 128;
    int bn = C;
    while (bn > 32 && pix_tiles * (d->nsrc * C / bn) < 96) bn >>= 1;
    static const bool no_split = getenv("VUB_DGRAD_NSPLIT") && getenv("VUB_DGRAD_NSPLIT")[0] == '0';
    if (!no_split) job.bn = bn;
  }
  job.n_split = d->nsrc == 2 ? C : 0;
  job.batch = d->batch;
  for (int s = 0; s < d->nsrc; ++s) {
    TileEpi& e = job.epi[s];
    e.addend = static_cast<const __nv_bfloat16*>(d->addend[s]);
    e.addend2 = static_cast<const __nv_bfloat16*>(d->addend2[s]);
    e.mask_elu = static_cast<const __nv_bfloat16*>(d->mask_elu[s]);
    e.out_raw = static_cast<__nv_bfloat16*>(d->dx[s]);
    e.out_h = d->hin;
    e.out_w = d->win;
    e.out_c = C;
    e.pix_stride = 1;
    e.n_valid = C;
    e.flags = d->flags & VUB_CONV_RELU;  // ReLU' instead of ELU' from the stored activation
  }

  {  // HBM-bound wide, thin layers: row-streaming kernel in data-gradient mode
    int rrc = 0;
    if (conv_rows_dgrad(d, stream, &rrc)) return rrc;
  }

  if (d->stride == 1 && !d2s && !(d->flags & VUB_CONV_RELU) && d->ksize == 3 && (d->nsrc == 1 || C % 128 == 0) &&
      conv_slab_shape_ok(cout_k, d->hin, d->win, d->nsrc * C, 9 * cout_k, ((long long)d->batch * d->hin * d->win + 127) / 128)) {
    // same shapes as the forward slab kernel with the channel roles swapped: A = dy, taps mirrored
    SlabJob sj;
    memset(&sj, 0, sizeof(sj));
    sj.x[0] = d->dy;
    sj.nsrc = 1;
    sj.c_src = cout_k;
    sj.batch = d->batch;
    sj.h = d->hin;
    sj.w = d->win;
    sj.w_packed = d->wd;
    sj.w_tiled = d->wd_tiled;
    sj.w_rows = d->nsrc * C;
    sj.ktot = (long long)taps * cout_k;
    sj.n_total = d->nsrc * C;
    sj.flip = 1;
    sj.n_split = d->nsrc == 2 ? C : 0;
    sj.epi[0] = job.epi[0];
    sj.epi[1] = job.epi[1];
    return conv_slab_launch(sj, stream);
  }

  if (d->stride == 1) {
    job.hout = d->hin;
    job.wout = d->win;
    const long long H = d->hin, W = d->win;
    if (!d2s) {
      job.n_src = 1;
      job.src[0] = TileSrc{d->dy, cout_k, d->win, d->hin, (long long)cout_k, W * cout_k, H * W * cout_k};
    } else {
      job.n_src = 4;
      for (int b1 = 0; b1 < 2; ++b1)
        for (int b2 = 0; b2 < 2; ++b2)
          job.src[b1 * 2 + b2] = TileSrc{static_cast<const char*>(d->dy) + ((long long)b1 * 2 * W + b2) * cmap * 2,
                                         cmap, d->win, d->hin, 2LL * cmap, 4LL * W * cmap, 4LL * H * W * cmap};
    }
    int ks = 0;
    for (int tap = 0; tap < taps; ++tap) {
      const int r = tap / d->ksize, q = tap % d->ksize;
      const int dh = taps == 9 ? 1 - r : 0, dw = taps == 9 ? 1 - q : 0;
      for (int m = 0; m < job.n_src; ++m)
        for (int cb = 0; cb < cblocks; ++cb, ++ks) {
          if (ks >= kTileMaxKSteps) return fail(VUB_E_SHAPE, "conv_dgrad: too many k-steps");
          job.ksteps[ks] = KStep{(uint8_t)m, (int8_t)dh, (int8_t)dw, (uint8_t)cb,
                                 (uint16_t)((tap * cout_k + m * cmap + cb * bkc) / bkc), 0};
        }
    }
    job.n_ksteps = ks;
    return conv_tiles_launch(job, stream);
  }

  // stride 2: y[oh] reads x[2*oh + r - 1]  =>  input row 2a + ph receives from taps r with (r - 1) % 2 == ph:
  //   ph = 0: r = 1 (oh = a);   ph = 1: r = 0 (oh = a + 1) and r = 2 (oh = a)
  job.hout = d->hin / 2;
  job.wout = d->win / 2;
  job.n_src = 1;
  {
    const long long Ho = d->hin / 2, Wo = d->win / 2;
    job.src[0] = TileSrc{d->dy, cout_k, (int)Wo, (int)Ho, (long long)cout_k, Wo * cout_k, Ho * Wo * cout_k};
  }
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      int ks = 0;
      for (int r = 0; r < 3; ++r) {
        if (((r + 1) & 1) != ph) continue;
        for (int q = 0; q < 3; ++q) {
          if (((q + 1) & 1) != pw) continue;
          const int dh = r == 0 ? 1 : 0, dw = q == 0 ? 1 : 0;
          for (int cb = 0; cb < cblocks; ++cb, ++ks)
            job.ksteps[ks] = KStep{0, (int8_t)dh, (int8_t)dw, (uint8_t)cb,
                                   (uint16_t)(((r * 3 + q) * cout_k + cb * bkc) / bkc), 0};
        }
      }
      job.n_ksteps = ks;
      job.epi[0].pix_stride = 2;
      job.epi[0].pix_off_h = ph;
      job.epi[0].pix_off_w = pw;
      const int rc = conv_tiles_launch(job, stream);
      if (rc) return rc;
    }
  return 0;
}
