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
      if (p.out_elu != nullptr) p.out_elu[o] = (p.flags & VUB_CONV_RELU) ? fmaxf(y, 0.f) : (y > 0.f ? y : expm1f(y));
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


// ------------------------------------------------------------------------------------------------
// fp32 validation path, training half (north_star: gradients within 1e-4 "in an fp32-accumulate validation mode").
// Same fused backward schedule as the bf16 path (vub_conv_dgrad / vub_conv_wgrad / vub_wn_bwd) with fp32
// activations, fp32 v / ||v|| matrices and fp32 FMA accumulation on CUDA cores.  Checkers, not fast paths.
// ------------------------------------------------------------------------------------------------
struct DgradF32Params {
  const float* dy;        // NHWC [batch, hout, wout, cout]; D2S: the shuffled tensor [batch, 2*hout, 2*wout, cout/4]
  const float* w;         // [K][w_pitch], K order (source, tap, channel), v / ||v||
  const float* scale;     // gamma * g per output channel
  float* dx[2];
  const float* mask_elu[2];
  const float* addend[2];
  const float* addend2[2];
  int nsrc, c_src, batch, hin, win, hout, wout, cout, w_pitch, ksize, stride;
  uint32_t flags;
  long long total;        // batch * hin * win * c_src elements per source
};

// element of dy for conv-output pixel (n, oh, ow), channel co
__device__ __forceinline__ long long dy_index_f32(int n, int oh, int ow, int co, int hout, int wout, int cout, bool d2s) {
  if (d2s) {
    const int cq = cout >> 2, qd = co / cq, cp = co - qd * cq;
    return (((long long)n * (2 * hout) + 2 * oh + (qd >> 1)) * (2 * wout) + 2 * ow + (qd & 1)) * cq + cp;
  }
  return (((long long)n * hout + oh) * wout + ow) * cout + co;
}

__global__ void __launch_bounds__(256) conv_dgrad_f32_kernel(const DgradF32Params p) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.total * p.nsrc) return;
  const int s = static_cast<int>(i / p.total);
  const long long e = i - (long long)s * p.total;
  const int ci = static_cast<int>(e % p.c_src);
  long long t = e / p.c_src;
  const int w = static_cast<int>(t % p.win);
  t /= p.win;
  const int h = static_cast<int>(t % p.hin);
  const int n = static_cast<int>(t / p.hin);
  const int pad = p.ksize / 2, taps = p.ksize * p.ksize;
  const bool d2s = (p.flags & VUB_CONV_D2S) != 0;
  float acc = 0.f;
  for (int tap = 0; tap < taps; ++tap) {
    int oh = h + pad - tap / p.ksize, ow = w + pad - tap % p.ksize;
    if (p.stride == 2) {
      if ((oh | ow) & 1) continue;
      oh >>= 1;
      ow >>= 1;
    }
    if (oh < 0 || oh >= p.hout || ow < 0 || ow >= p.wout) continue;
    const float* wrow = p.w + (long long)((s * taps + tap) * p.c_src + ci) * p.w_pitch;
    if (!d2s) {
      const float* dyp = p.dy + dy_index_f32(n, oh, ow, 0, p.hout, p.wout, p.cout, false);
      for (int co = 0; co < p.cout; ++co) acc = fmaf(__ldg(dyp + co) * __ldg(p.scale + co), __ldg(wrow + co), acc);
    } else {
      for (int co = 0; co < p.cout; ++co)
        acc = fmaf(__ldg(p.dy + dy_index_f32(n, oh, ow, co, p.hout, p.wout, p.cout, true)) * __ldg(p.scale + co),
                   __ldg(wrow + co), acc);
    }
  }
  if (p.mask_elu[s] != nullptr) {
    const float ev = p.mask_elu[s][e];
    acc *= ev > 0.f ? 1.0f : ((p.flags & VUB_CONV_RELU) ? 0.0f : ev + 1.0f);
  }
  if (p.addend[s] != nullptr) acc += p.addend[s][e];
  if (p.addend2[s] != nullptr) acc += p.addend2[s][e];
  p.dx[s][e] = acc;
}

struct WgradF32Params {
  const float* dy;
  const float* x[2];
  float* gt;       // [n_split][cout rows, pitch gt_pitch][K]
  float* dy_sum;   // [n_split][cout]
  int nsrc, c_src, batch, hin, win, hout, wout, cout, ksize, stride, ktot;
  uint32_t flags;
  int n_split;
  long long split_stride, sum_stride, total_pix;
  int gt_pitch;
};

// gt[co][k] = sum_pix dy[pix, co] * x_patch[pix, k]; 32 x 32 tile per block, pixels of the block's split streamed
// through shared memory 32 at a time; blockIdx = (k tile, co tile, split).
__global__ void __launch_bounds__(256) conv_wgrad_f32_kernel(const WgradF32Params p) {
  __shared__ float dyS[32][33];
  __shared__ float xS[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7
  const int k0 = blockIdx.x * 32, co0 = blockIdx.y * 32, split = blockIdx.z;
  const long long per = (p.total_pix + p.n_split - 1) / p.n_split;
  const long long p_begin = per * split;
  const long long p_end = p_begin + per < p.total_pix ? p_begin + per : p.total_pix;
  const int pad = p.ksize / 2, taps = p.ksize * p.ksize;
  const bool d2s = (p.flags & VUB_CONV_D2S) != 0;
  // this thread's k column (for the x loads): decode (source, tap, channel) once
  const int k = k0 + tx;
  const bool k_ok = k < p.ktot;
  const int ks = k_ok ? k / (taps * p.c_src) : 0;
  const int ktap = k_ok ? (k / p.c_src) % taps : 0;
  const int kci = k_ok ? k % p.c_src : 0;
  const int kh = ktap / p.ksize, kw = ktap % p.ksize;
  const int co_l = co0 + tx;  // this thread's co column for the dy loads
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long base = p_begin; base < p_end; base += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int px = ty + 8 * i;
      const long long pix = base + px;
      float dv = 0.f, xv = 0.f;
      if (pix < p_end) {
        const int ow = static_cast<int>(pix % p.wout);
        const long long t = pix / p.wout;
        const int oh = static_cast<int>(t % p.hout);
        const int n = static_cast<int>(t / p.hout);
        if (co_l < p.cout) dv = __ldg(p.dy + dy_index_f32(n, oh, ow, co_l, p.hout, p.wout, p.cout, d2s));
        if (k_ok) {
          const int ih = oh * p.stride + kh - pad, iw = ow * p.stride + kw - pad;
          if (ih >= 0 && ih < p.hin && iw >= 0 && iw < p.win)
            xv = __ldg(p.x[ks] + (((long long)n * p.hin + ih) * p.win + iw) * p.c_src + kci);
        }
      }
      dyS[px][tx] = dv;
      xS[px][tx] = xv;
    }
    __syncthreads();
#pragma unroll 8
    for (int px = 0; px < 32; ++px) {
      const float xv = xS[px][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dv = dyS[px][ty + 8 * i];
        acc[i] = fmaf(dv, xv, acc[i]);
        sum[i] += dv;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty + 8 * i;
    if (co >= p.cout) continue;
    if (k_ok) p.gt[split * p.split_stride + (long long)co * p.gt_pitch + k] = acc[i];
    if (blockIdx.x == 0 && tx == 0) p.dy_sum[split * p.sum_stride + co] = sum[i];
  }
}

// latent head backward in fp32 (same math as latent_bwd_kernel, fp32 NHWC dz in / d_ml out); thread per element
__global__ void __launch_bounds__(256)
    latent_bwd_f32_kernel(const float* __restrict__ ls, const float* __restrict__ eps, const float* __restrict__ d_mu,
                          const float* __restrict__ d_ls, const float* __restrict__ dz, int c, int hw, long long total,
                          float* __restrict__ d_ml) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  const long long pix = i / c;
  const long long n = pix / hw, sp = pix - n * hw;
  const long long o = (n * c + ch) * hw + sp;  // NCHW
  const float dzv = dz != nullptr ? dz[i] : 0.f;
  const float l = ls[o];
  const float e = eps != nullptr ? eps[o] : 0.f;
  const float dm = (d_mu != nullptr ? d_mu[o] : 0.f) + dzv;
  const float dl = (d_ls != nullptr ? d_ls[o] : 0.f) + dzv * e * expf(l);
  d_ml[pix * (2 * c) + ch] = dm;
  d_ml[pix * (2 * c) + c + ch] = dl * l * (1.0f - l);
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

extern "C" int vub_conv_dgrad_f32(const vub_dgrad_f32_desc* d, void* stream) {
  clear_error();
  if (!d) return fail(VUB_E_NULL, "conv_dgrad_f32: desc is NULL");
  if (!d->dy || !d->w || !d->scale || !d->dx[0]) return fail(VUB_E_NULL, "conv_dgrad_f32: dy/w/scale/dx NULL");
  if (d->nsrc != 1 && d->nsrc != 2) return fail(VUB_E_SHAPE, "conv_dgrad_f32: nsrc must be 1 or 2");
  if (d->nsrc == 2 && !d->dx[1]) return fail(VUB_E_NULL, "conv_dgrad_f32: dx[1] NULL with nsrc == 2");
  if ((d->ksize != 1 && d->ksize != 3) || (d->stride != 1 && d->stride != 2))
    return fail(VUB_E_SHAPE, "conv_dgrad_f32: ksize in {1,3}, stride in {1,2}");
  if (d->batch <= 0 || d->hin <= 0 || d->win <= 0 || d->c_src <= 0 || d->cout <= 0 || d->hin % d->stride ||
      d->win % d->stride || d->w_pitch < d->cout)
    return fail(VUB_E_SHAPE, "conv_dgrad_f32: bad extents");
  if ((d->flags & VUB_CONV_D2S) && (d->cout % 4 || d->stride != 1))
    return fail(VUB_E_FLAGS, "conv_dgrad_f32: D2S needs cout %% 4 == 0 and stride 1");
  DgradF32Params p;
  p.dy = d->dy;
  p.w = d->w;
  p.scale = d->scale;
  for (int s = 0; s < 2; ++s) {
    p.dx[s] = d->dx[s];
    p.mask_elu[s] = d->mask_elu[s];
    p.addend[s] = d->addend[s];
    p.addend2[s] = d->addend2[s];
  }
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
  p.total = (long long)d->batch * d->hin * d->win * d->c_src;
  const long long n = p.total * p.nsrc;
  conv_dgrad_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("conv_dgrad_f32");
}

extern "C" int vub_conv_wgrad_f32(const vub_wgrad_f32_desc* d, void* stream) {
  clear_error();
  if (!d) return fail(VUB_E_NULL, "conv_wgrad_f32: desc is NULL");
  if (!d->dy || !d->x[0] || !d->gt || !d->dy_sum) return fail(VUB_E_NULL, "conv_wgrad_f32: dy/x/gt/dy_sum NULL");
  if (d->nsrc != 1 && d->nsrc != 2) return fail(VUB_E_SHAPE, "conv_wgrad_f32: nsrc must be 1 or 2");
  if (d->nsrc == 2 && !d->x[1]) return fail(VUB_E_NULL, "conv_wgrad_f32: x[1] NULL with nsrc == 2");
  if ((d->ksize != 1 && d->ksize != 3) || (d->stride != 1 && d->stride != 2))
    return fail(VUB_E_SHAPE, "conv_wgrad_f32: ksize in {1,3}, stride in {1,2}");
  if (d->batch <= 0 || d->hin <= 0 || d->win <= 0 || d->c_src <= 0 || d->cout <= 0 || d->hin % d->stride ||
      d->win % d->stride || d->n_split <= 0 || d->n_split > 65535)
    return fail(VUB_E_SHAPE, "conv_wgrad_f32: bad extents");
  WgradF32Params p;
  p.dy = d->dy;
  p.x[0] = d->x[0];
  p.x[1] = d->x[1];
  p.gt = d->gt;
  p.dy_sum = d->dy_sum;
  p.nsrc = d->nsrc;
  p.c_src = d->c_src;
  p.batch = d->batch;
  p.hin = d->hin;
  p.win = d->win;
  p.hout = d->hin / d->stride;
  p.wout = d->win / d->stride;
  p.cout = d->cout;
  p.ksize = d->ksize;
  p.stride = d->stride;
  p.ktot = d->nsrc * d->ksize * d->ksize * d->c_src;
  p.flags = d->flags;
  p.n_split = d->n_split;
  p.split_stride = d->split_stride;
  p.sum_stride = d->sum_stride;
  p.gt_pitch = d->gt_pitch;
  p.total_pix = (long long)d->batch * p.hout * p.wout;
  if (p.gt_pitch < p.ktot) return fail(VUB_E_SHAPE, "conv_wgrad_f32: gt_pitch %d < K %d", p.gt_pitch, p.ktot);
  dim3 grid((p.ktot + 31) / 32, (p.cout + 31) / 32, p.n_split);
  conv_wgrad_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("conv_wgrad_f32");
}

extern "C" int vub_latent_bwd_f32(const float* logstd_nchw, const float* eps_nchw, const float* d_mu_nchw,
                                  const float* d_ls_nchw, const float* dz_nhwc, int32_t batch, int32_t c, int32_t h,
                                  int32_t w, float* d_ml_nhwc, void* stream) {
  clear_error();
  if (!logstd_nchw || !d_ml_nhwc) return fail(VUB_E_NULL, "latent_bwd_f32: NULL pointer");
  if (batch <= 0 || c <= 0 || h <= 0 || w <= 0) return fail(VUB_E_SHAPE, "latent_bwd_f32: bad shape");
  const long long total = (long long)batch * h * w * c;
  latent_bwd_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logstd_nchw, eps_nchw, d_mu_nchw, d_ls_nchw, dz_nhwc, c, h * w, total, d_ml_nhwc);
  return check_launch("latent_bwd_f32");
}
