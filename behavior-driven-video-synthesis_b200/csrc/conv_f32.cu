// fp32 validation path (north_star: "1e-4 in an fp32-accumulate validation mode"): the same fused conv
// (NormConv2d + VunetRNB elementwise work + DepthToSpace + Downsample) with fp32 operands, fp32 FMA
// accumulation and fp32 activations on CUDA cores.  Direct convolution, no tensor cores, not a fast path:
// it exists so that the fused SCHEDULE (operand packing order, concat elimination, raw/ELU pairs,
// DepthToSpace addressing, latent head) can be checked against the reference to fp32 round-off, separating
// schedule errors from bf16 rounding.  Replaces lib/modules.py:140-145, 221-233, 24-34, 148-182.
#include "common.h"

namespace vub {

constexpr int kF32Threads = 128;
constexpr int kF32Pix = 8;        // output pixels per block
constexpr int kF32Quads = 16;     // channel quads per pass (64 output channels)

struct ConvF32Params {
  const float* x[2];
  const float* w;       // [K][w_pitch], K order (source, tap, channel), v / ||v||
  const float* scale;
  const float* shift;
  const float* residual;
  float* out_raw;
  float* out_elu;
  int nsrc, c_src, batch, hin, win, hout, wout, cout, w_pitch, ksize, stride;
  uint32_t flags;
  long long total_pix;
};

__global__ void __launch_bounds__(kF32Threads) conv_f32_kernel(const ConvF32Params p) {
  const int quad = threadIdx.x % kF32Quads;
  const long long pix = (long long)blockIdx.x * kF32Pix + threadIdx.x / kF32Quads;
  if (pix >= p.total_pix) return;
  const int ow = static_cast<int>(pix % p.wout);
  const long long t = pix / p.wout;
  const int oh = static_cast<int>(t % p.hout);
  const int n = static_cast<int>(t / p.hout);
  const int pad = p.ksize / 2, taps = p.ksize * p.ksize;
  for (int co = quad * 4; co < p.cout; co += kF32Quads * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < p.nsrc; ++s)
      for (int tap = 0; tap < taps; ++tap) {
        const int ih = oh * p.stride + tap / p.ksize - pad, iw = ow * p.stride + tap % p.ksize - pad;
        if (ih < 0 || ih >= p.hin || iw < 0 || iw >= p.win) continue;
        const float* xp = p.x[s] + (((long long)n * p.hin + ih) * p.win + iw) * p.c_src;
        const float* wp = p.w + (long long)((s * taps + tap) * p.c_src) * p.w_pitch + co;
        for (int c = 0; c < p.c_src; ++c) {
          const float xv = __ldg(xp + c);
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + (long long)c * p.w_pitch));
          acc[0] = fmaf(xv, w4.x, acc[0]);
          acc[1] = fmaf(xv, w4.y, acc[1]);
          acc[2] = fmaf(xv, w4.z, acc[2]);
          acc[3] = fmaf(xv, w4.w, acc[3]);
        }
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = co + i;
      if (c >= p.cout) break;
      float y = fmaf(acc[i], p.scale[c], p.shift[c]);
      long long o;
      if (p.flags & VUB_CONV_F32_NCHW) {
        o = (((long long)n * p.cout + c) * p.hout + oh) * p.wout + ow;
      } else if (p.flags & VUB_CONV_D2S) {
        const int cq = p.cout >> 2, qd = c / cq, cp = c - qd * cq;
        o = (((long long)n * (2 * p.hout) + 2 * oh + (qd >> 1)) * (2 * p.wout) + 2 * ow + (qd & 1)) * cq + cp;
      } else {
        o = pix * p.cout + c;
      }
      if (p.residual != nullptr) y += p.residual[o];
      if (p.out_raw != nullptr) p.out_raw[o] = y;
      if (p.out_elu != nullptr) p.out_elu[o] = y > 0.f ? y : expm1f(y);
    }
  }
}

// NCHW fp32 -> NHWC fp32 (raw and/or exact ELU); one thread per element of the output.
__global__ void __launch_bounds__(256)
    nchw_to_nhwc_f32_kernel(const float* __restrict__ src, int c, long long hw, long long total,
                            float* __restrict__ raw, float* __restrict__ elu) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  const long long pix = i / c;
  const long long n = pix / hw, sp = pix - n * hw;
  const float v = src[(n * c + ch) * hw + sp];
  if (raw != nullptr) raw[i] = v;
  if (elu != nullptr) elu[i] = v > 0.f ? v : expm1f(v);
}

}  // namespace vub

using namespace vub;

extern "C" int vub_conv_fwd_f32(const vub_conv_f32_desc* d, void* stream) {
  clear_error();
  if (!d) return fail(VUB_E_NULL, "conv_fwd_f32: desc is NULL");
  if (!d->x[0] || !d->w || !d->scale || !d->shift) return fail(VUB_E_NULL, "conv_fwd_f32: x/w/scale/shift NULL");
  if (d->nsrc != 1 && d->nsrc != 2) return fail(VUB_E_SHAPE, "conv_fwd_f32: nsrc must be 1 or 2");
  if (d->nsrc == 2 && !d->x[1]) return fail(VUB_E_NULL, "conv_fwd_f32: x[1] NULL with nsrc == 2");
  if (d->ksize != 1 && d->ksize != 3) return fail(VUB_E_SHAPE, "conv_fwd_f32: ksize must be 1 or 3");
  if (d->stride != 1 && d->stride != 2) return fail(VUB_E_SHAPE, "conv_fwd_f32: stride must be 1 or 2");
  if (d->batch <= 0 || d->hin <= 0 || d->win <= 0 || d->c_src <= 0 || d->cout <= 0 || d->hin % d->stride ||
      d->win % d->stride)
    return fail(VUB_E_SHAPE, "conv_fwd_f32: bad extents");
  if (d->w_pitch < ((d->cout + 3) & ~3) || d->w_pitch % 4 || !aligned16(d->w))
    return fail(VUB_E_SHAPE, "conv_fwd_f32: w_pitch %d must be a multiple of 4 >= cout %d (w 16B aligned)", d->w_pitch,
                d->cout);
  if (!d->out_raw && !d->out_elu) return fail(VUB_E_FLAGS, "conv_fwd_f32: no output");
  if ((d->flags & VUB_CONV_D2S) && (d->cout % 4 || d->residual || (d->flags & VUB_CONV_F32_NCHW)))
    return fail(VUB_E_FLAGS, "conv_fwd_f32: D2S needs cout % 4 == 0, no residual, NHWC output");
  ConvF32Params p;
  p.x[0] = d->x[0];
  p.x[1] = d->x[1];
  p.w = d->w;
  p.scale = d->scale;
  p.shift = d->shift;
  p.residual = d->residual;
  p.out_raw = d->out_raw;
  p.out_elu = d->out_elu;
  p.nsrc = d->nsrc;
  p.c_src = d->c_src;
  p.batch = d->batch;
  p.hin = d->hin;
  p.win = d->win;
  p.hout = d->hin / d->stride;
  p.wout = d->win / d->stride;
  p.cout = d->cout;
  p.w_pitch = d->w_pitch;
  p.ksize = d->ksize;
  p.stride = d->stride;
  p.flags = d->flags;
  p.total_pix = (long long)d->batch * p.hout * p.wout;
  const unsigned grid = static_cast<unsigned>((p.total_pix + kF32Pix - 1) / kF32Pix);
  conv_f32_kernel<<<grid, kF32Threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("conv_fwd_f32");
}

extern "C" int vub_nchw_f32_to_nhwc_f32(const float* src, int32_t batch, int32_t c, int32_t h, int32_t w,
                                        float* dst_raw, float* dst_elu, void* stream) {
  clear_error();
  if (!src || (!dst_raw && !dst_elu)) return fail(VUB_E_NULL, "nchw_to_nhwc_f32: NULL pointer");
  if (batch <= 0 || c <= 0 || h <= 0 || w <= 0) return fail(VUB_E_SHAPE, "nchw_to_nhwc_f32: bad shape");
  const long long hw = (long long)h * w, total = hw * batch * c;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  nchw_to_nhwc_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, c, hw, total, dst_raw, dst_elu);
  return check_launch("nchw_to_nhwc_f32");
}
