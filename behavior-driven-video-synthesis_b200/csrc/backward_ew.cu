// Memory-bound kernels of the training half of the VUnet path: parameter gradients of all
// weight-normalised convs from their weight correlations (one launch per step), the stem's weight
// correlation (K = 3/30 input channels: CUDA cores), the latent head's backward and the image-gradient
// layout converter.  Coalesced accesses, warp-shuffle / shared-memory reductions, fp32 arithmetic.
//
// Replaces what autograd runs in the reference: aten::_weight_norm_interface_backward and the mul/add
// backward of lib/modules.py:135-145, sigmoid/exp/mul/add backward of models/vunets.py:575,594-597.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace vub {

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two block-wide sums at once (blockDim.x multiple of 32, <= 1024); every thread gets both.
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* red /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum_b(a);
  b = warp_sum_b(b);
  if (lane == 0) red[warp] = make_float2(a, b);
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float2 t = (threadIdx.x < nw) ? red[threadIdx.x] : make_float2(0.f, 0.f);
  if (warp == 0) {
    t.x = warp_sum_b(t.x);
    t.y = warp_sum_b(t.y);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------------------------------------
// Weight-norm / affine parameter gradients: one block per output channel of any layer.
// ------------------------------------------------------------------------------------------------
constexpr int kWnbThreads = 128;
constexpr int kWnbDefaultLen = 2304;  // cin * k * k of the largest VunetAlter layer (256 x 3 x 3)
constexpr int kWnbMaxLen = 9216;      // ... of the largest VunetOrg layer (1024 x 3 x 3): 36 KB of dynamic shared memory

__global__ void __launch_bounds__(kWnbThreads)
    wn_bwd_kernel(const vub_wnb_entry* __restrict__ table, int n_entries, int accumulate) {
  __shared__ float2 red[32];
  extern __shared__ float g_s[];  // this channel's correlation, summed over the pixel splits, in weight order (max_len floats)
  const int row = blockIdx.x;
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].row_start <= row) lo = mid;
    else hi = mid - 1;
  }
  const vub_wnb_entry e = table[lo];
  const int co = row - e.row_start;
  if (co >= e.cout) return;
  const int len = e.cin * e.ktaps;
  const float* v = e.v + (size_t)co * len;
  const float* grow = e.gt + (size_t)(e.gt_col0 + co) * e.gt_pitch;
  // sum the pixel splits (coalesced along k), stored at the weight index i = (ci, tap) of column
  // k = (src*ktaps + tap)*c_src + cs
  auto weight_index = [&](int k) {
    int cs, t;
    if ((e.c_src & (e.c_src - 1)) == 0) {  // convs: power-of-two channels per source -> shifts instead of divisions
      const int sh = __ffs(e.c_src) - 1;
      cs = k & (e.c_src - 1);
      t = k >> sh;
    } else {
      cs = k % e.c_src;
      t = k / e.c_src;
    }
    int tap, src;
    if (e.ktaps == 9) {
      src = t / 9;
      tap = t - 9 * src;
    } else {
      src = t / e.ktaps;
      tap = t - src * e.ktaps;
    }
    return (src * e.c_src + cs) * e.ktaps + tap;
  };
  const bool vec = (len & 3) == 0 && (e.gt_pitch & 3) == 0 && (e.split_stride & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(grow) & 15) == 0;
  if (vec) {
    // 16-byte loads, two independent chains: 8 x 16 B in flight per thread (the kernel is a pure HBM stream over the
    // splits' partials, ~1.2 GB per step)
    for (int k = threadIdx.x * 4; k < len; k += kWnbThreads * 4) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a, d = a;
      int sidx = 0;
      for (; sidx + 4 <= e.n_split; sidx += 4) {
        const float* q = grow + (size_t)sidx * e.split_stride + k;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(q));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(q + e.split_stride));
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(q + 2 * e.split_stride));
        const float4 v3 = __ldg(reinterpret_cast<const float4*>(q + 3 * e.split_stride));
        a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
        b.x += v1.x; b.y += v1.y; b.z += v1.z; b.w += v1.w;
        c.x += v2.x; c.y += v2.y; c.z += v2.z; c.w += v2.w;
        d.x += v3.x; d.y += v3.y; d.z += v3.z; d.w += v3.w;
      }
      for (; sidx < e.n_split; ++sidx) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(grow + (size_t)sidx * e.split_stride + k));
        a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
      }
      // same summation order as the scalar path: (s0 + s1) + (s2 + s3) per element
      g_s[weight_index(k + 0)] = (a.x + b.x) + (c.x + d.x);
      g_s[weight_index(k + 1)] = (a.y + b.y) + (c.y + d.y);
      g_s[weight_index(k + 2)] = (a.z + b.z) + (c.z + d.z);
      g_s[weight_index(k + 3)] = (a.w + b.w) + (c.w + d.w);
    }
  } else {
    for (int k = threadIdx.x; k < len; k += kWnbThreads) {
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;  // independent chains: keep 4+ loads in flight per thread
      int sidx = 0;
      for (; sidx + 4 <= e.n_split; sidx += 4) {
        const float* q = grow + (size_t)sidx * e.split_stride + k;
        g0 += __ldg(q);
        g1 += __ldg(q + e.split_stride);
        g2 += __ldg(q + 2 * e.split_stride);
        g3 += __ldg(q + 3 * e.split_stride);
      }
      for (; sidx < e.n_split; ++sidx) g0 += __ldg(grow + (size_t)sidx * e.split_stride + k);
      g_s[weight_index(k)] = (g0 + g1) + (g2 + g3);
    }
  }
  __syncthreads();
  auto g_at = [&](int i) { return g_s[i]; };
  float ss = 0.f, vg = 0.f;
  for (int i = threadIdx.x; i < len; i += kWnbThreads) {
    const float x = v[i];
    ss = fmaf(x, x, ss);
    vg = fmaf(x, g_at(i), vg);
  }
  const float2 r = block_sum2(ss, vg, red);
  const float inv = 1.0f / sqrtf(r.x);
  const float dot = r.y * inv;  // v_hat . G
  const float ga = e.gamma[co], gg = e.g[co], bi = e.bias[co];
  float S = 0.f;
  for (int sidx = 0; sidx < e.n_split; ++sidx) S += e.dy_sum[(size_t)sidx * e.sum_stride + e.gt_col0 + co];
  const float k = ga * gg * inv;
  float* dv = e.d_v + (size_t)co * len;
  for (int i = threadIdx.x; i < len; i += kWnbThreads) {
    const float val = k * (g_at(i) - v[i] * inv * dot);
    dv[i] = accumulate ? dv[i] + val : val;
  }
  if (threadIdx.x == 0) {
    const float dg = ga * dot, dgam = fmaf(gg, dot, bi * S), dbeta = S, dbias = ga * S;
    if (accumulate) {
      e.d_g[co] += dg;
      e.d_gamma[co] += dgam;
      e.d_beta[co] += dbeta;
      e.d_bias[co] += dbias;
    } else {
      e.d_g[co] = dg;
      e.d_gamma[co] = dgam;
      e.d_beta[co] = dbeta;
      e.d_bias[co] = dbias;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stem correlation: gt[co][ci] += sum_pix x[n, ci, pix] * dy[n, pix, co];  dy_sum[co] += sum dy.
// One warp per 32-pixel chunk: lane l holds x[ci][p0 + l] for every input channel (coalesced 128-byte
// loads per channel plane) and owns output channel(s) l (COUT/32 per lane); the chunk's dy rows
// (32 x COUT bf16, contiguous) are staged in shared memory with 16-byte loads; per pixel the x values are
// broadcast with warp shuffles.  Accumulators live in registers over the warp's whole pixel range; one
// shared-memory reduction per block and one red.global.add per (ci, co) and block at the end.
// ------------------------------------------------------------------------------------------------
constexpr int kSbThreads = 256;
constexpr int kSbWarps = kSbThreads / 32;
constexpr int kSbPix = 32;
constexpr int kSbMaxCx = 32;

template <int COUT, int CXMAX>
__global__ void __launch_bounds__(kSbThreads)
    stem_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int cx, int hw,
                    long long n_chunks, float* __restrict__ gt, float* __restrict__ dy_sum) {
  constexpr int PER = COUT / 32;
  __shared__ __align__(16) __nv_bfloat16 ds[kSbWarps][kSbPix * COUT];
  __shared__ float red[CXMAX + 1][COUT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[CXMAX][PER];
  float sdy[PER];
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    sdy[t] = 0.f;
#pragma unroll
    for (int ci = 0; ci < CXMAX; ++ci) acc[ci][t] = 0.f;
  }
  for (int i = threadIdx.x; i < (CXMAX + 1) * COUT; i += kSbThreads) (&red[0][0])[i] = 0.f;
  const int chunks_per_img = hw / kSbPix;
  for (long long ch = (long long)blockIdx.x * kSbWarps + warp; ch < n_chunks; ch += (long long)gridDim.x * kSbWarps) {
    const long long n = ch / chunks_per_img;
    const int p0 = static_cast<int>(ch - n * chunks_per_img) * kSbPix;
    float xr[CXMAX];
#pragma unroll
    for (int ci = 0; ci < CXMAX; ++ci) xr[ci] = ci < cx ? __ldg(x + (n * cx + ci) * hw + p0 + lane) : 0.f;
    const uint4* src = reinterpret_cast<const uint4*>(dy + (n * hw + p0) * COUT);
    uint4* dst = reinterpret_cast<uint4*>(ds[warp]);
#pragma unroll
    for (int k = 0; k < COUT / 8; ++k) dst[lane + 32 * k] = __ldg(src + lane + 32 * k);
    __syncwarp();
#pragma unroll 4
    for (int j = 0; j < kSbPix; ++j) {
      float d[PER];
#pragma unroll
      for (int t = 0; t < PER; ++t) {
        d[t] = __bfloat162float(ds[warp][j * COUT + lane * PER + t]);
        sdy[t] += d[t];
      }
#pragma unroll
      for (int ci = 0; ci < CXMAX; ++ci) {
        const float xv = __shfl_sync(0xffffffffu, xr[ci], j);
#pragma unroll
        for (int t = 0; t < PER; ++t) acc[ci][t] = fmaf(xv, d[t], acc[ci][t]);
      }
    }
    __syncwarp();
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < PER; ++t) {
#pragma unroll
    for (int ci = 0; ci < CXMAX; ++ci)
      if (ci < cx) atomicAdd(&red[ci][lane * PER + t], acc[ci][t]);
    atomicAdd(&red[CXMAX][lane * PER + t], sdy[t]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cx * COUT; i += kSbThreads) atomicAdd(gt + i, red[i % cx][i / cx]);  // gt[co][ci]
  for (int i = threadIdx.x; i < COUT; i += kSbThreads) atomicAdd(dy_sum + i, red[CXMAX][i]);
}

// ------------------------------------------------------------------------------------------------
// Latent head backward: one thread per 4 channels of one pixel (mirror of latent_kernel).
// ------------------------------------------------------------------------------------------------
constexpr int kLbThreads = 128;

__global__ void __launch_bounds__(kLbThreads)
    latent_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ eps,
                      const float* __restrict__ d_mu, const float* __restrict__ d_ls,
                      const __nv_bfloat16* __restrict__ dz, int c, int hw, long long total4,
                      __nv_bfloat16* __restrict__ d_ml) {
  const long long idx = (long long)blockIdx.x * kLbThreads + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = c >> 2;
  const long long pix = idx / c4;
  const int ch = static_cast<int>(idx - pix * c4) * 4;
  const long long n = pix / hw;
  const int sp = static_cast<int>(pix - n * hw);
  float dzv[4] = {0.f, 0.f, 0.f, 0.f};
  if (dz != nullptr) {
    const uint2 u = *reinterpret_cast<const uint2*>(dz + pix * c + ch);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    dzv[0] = __low2float(a); dzv[1] = __high2float(a); dzv[2] = __low2float(b); dzv[3] = __high2float(b);
  }
  float gm[4], gl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long o = (n * c + ch + i) * hw + sp;  // NCHW
    const float l = ls[o];
    const float e = eps != nullptr ? eps[o] : 0.f;
    const float dm = (d_mu != nullptr ? d_mu[o] : 0.f) + dzv[i];
    const float dl = (d_ls != nullptr ? d_ls[o] : 0.f) + dzv[i] * e * expf(l);
    gm[i] = dm;
    gl[i] = dl * l * (1.0f - l);
  }
  (void)mu;
  __nv_bfloat162 a = __floats2bfloat162_rn(gm[0], gm[1]), b = __floats2bfloat162_rn(gm[2], gm[3]);
  *reinterpret_cast<uint2*>(d_ml + pix * (2 * c) + ch) =
      make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  a = __floats2bfloat162_rn(gl[0], gl[1]);
  b = __floats2bfloat162_rn(gl[2], gl[3]);
  *reinterpret_cast<uint2*>(d_ml + pix * (2 * c) + c + ch) =
      make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// ------------------------------------------------------------------------------------------------
// NCHW fp32 [B, c, hw] -> NHWC bf16 [B, hw, c_pad] with zero-filled pad channels; one thread per pixel.
// ------------------------------------------------------------------------------------------------
template <int CPAD>
__global__ void __launch_bounds__(256)
    nchw_to_nhwc_pad_kernel(const float* __restrict__ src, int c, long long hw, long long total,
                            __nv_bfloat16* __restrict__ dst) {
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= total) return;
  const long long n = pix / hw, sp = pix - n * hw;
  float v[CPAD];
#pragma unroll
  for (int i = 0; i < CPAD; ++i) v[i] = i < c ? __ldg(src + (n * c + i) * hw + sp) : 0.f;
  uint4* o = reinterpret_cast<uint4*>(dst + pix * CPAD);
#pragma unroll
  for (int j = 0; j < CPAD / 8; ++j) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
    __nv_bfloat162 cc = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
    o[j] = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b),
                      *reinterpret_cast<uint32_t*>(&cc), *reinterpret_cast<uint32_t*>(&d));
  }
}

}  // namespace vub

using namespace vub;

extern "C" int vub_wn_bwd(const vub_wnb_entry* table_dev, int32_t n_entries, int32_t total_rows,
                          int32_t accumulate, int32_t max_len, void* stream) {
  clear_error();
  if (!table_dev) return fail(VUB_E_NULL, "wn_bwd: table is NULL");
  if (n_entries <= 0 || total_rows <= 0) return fail(VUB_E_SHAPE, "wn_bwd: empty table");
  if (max_len <= 0) max_len = kWnbDefaultLen;
  if (max_len > kWnbMaxLen) return fail(VUB_E_SHAPE, "wn_bwd: max_len %d > %d", max_len, kWnbMaxLen);
  wn_bwd_kernel<<<total_rows, kWnbThreads, (size_t)max_len * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      table_dev, n_entries, accumulate);
  return check_launch("wn_bwd");
}

extern "C" int vub_stem_bwd(const float* x_nchw, const void* dy_nhwc, int32_t batch, int32_t cx, int32_t h,
                            int32_t w, int32_t cout, float* gt, float* dy_sum, void* stream) {
  clear_error();
  if (!x_nchw || !dy_nhwc || !gt || !dy_sum) return fail(VUB_E_NULL, "stem_bwd: NULL pointer");
  const long long hw = (long long)h * w;
  if (batch <= 0 || cx <= 0 || cx > kSbMaxCx || hw <= 0 || hw % kSbPix || hw > (1 << 30))
    return fail(VUB_E_SHAPE, "stem_bwd: batch %d cx %d h %d w %d (cx <= %d, h*w %% %d == 0)", batch, cx, h, w,
                kSbMaxCx, kSbPix);
  const long long n_chunks = (long long)batch * (hw / kSbPix);
  const long long blocks = (n_chunks + kSbWarps - 1) / kSbWarps;
  const unsigned grid = static_cast<unsigned>(blocks < sm_count() * 4 ? blocks : sm_count() * 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* dyp = static_cast<const __nv_bfloat16*>(dy_nhwc);
  const int hwi = static_cast<int>(hw);
  if (cout == 32 && cx <= 4)
    stem_bwd_kernel<32, 4><<<grid, kSbThreads, 0, s>>>(x_nchw, dyp, cx, hwi, n_chunks, gt, dy_sum);
  else if (cout == 32)
    stem_bwd_kernel<32, kSbMaxCx><<<grid, kSbThreads, 0, s>>>(x_nchw, dyp, cx, hwi, n_chunks, gt, dy_sum);
  else if (cout == 64 && cx <= 4)
    stem_bwd_kernel<64, 4><<<grid, kSbThreads, 0, s>>>(x_nchw, dyp, cx, hwi, n_chunks, gt, dy_sum);
  else if (cout == 64)
    stem_bwd_kernel<64, kSbMaxCx><<<grid, kSbThreads, 0, s>>>(x_nchw, dyp, cx, hwi, n_chunks, gt, dy_sum);
  else
    return fail(VUB_E_SHAPE, "stem_bwd: cout %d not in {32,64}", cout);
  return check_launch("stem_bwd");
}

extern "C" int vub_latent_bwd(const float* mu_nchw, const float* logstd_nchw, const float* eps_nchw,
                              const float* d_mu_nchw, const float* d_ls_nchw, const void* dz_nhwc, int32_t batch,
                              int32_t c, int32_t h, int32_t w, void* d_ml_nhwc, void* stream) {
  clear_error();
  if (!logstd_nchw || !d_ml_nhwc) return fail(VUB_E_NULL, "latent_bwd: NULL pointer");
  if (batch <= 0 || c <= 0 || (c & 3) || h <= 0 || w <= 0)
    return fail(VUB_E_SHAPE, "latent_bwd: batch %d c %d h %d w %d (c %% 4 == 0)", batch, c, h, w);
  const long long total4 = (long long)batch * h * w * (c / 4);
  const unsigned grid = static_cast<unsigned>((total4 + kLbThreads - 1) / kLbThreads);
  latent_bwd_kernel<<<grid, kLbThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      mu_nchw, logstd_nchw, eps_nchw, d_mu_nchw, d_ls_nchw, static_cast<const __nv_bfloat16*>(dz_nhwc), c, h * w,
      total4, static_cast<__nv_bfloat16*>(d_ml_nhwc));
  return check_launch("latent_bwd");
}

extern "C" int vub_nchw_f32_to_nhwc_bf16_pad(const float* src, int32_t batch, int32_t c, int32_t h, int32_t w,
                                             int32_t c_pad, void* dst, void* stream) {
  clear_error();
  if (!src || !dst) return fail(VUB_E_NULL, "nchw_to_nhwc_pad: NULL pointer");
  if (batch <= 0 || c <= 0 || h <= 0 || w <= 0 || c > c_pad) return fail(VUB_E_SHAPE, "nchw_to_nhwc_pad: bad shape");
  if (!aligned16(dst)) return fail(VUB_E_ALIGN, "nchw_to_nhwc_pad: dst not 16B aligned");
  const long long hw = (long long)h * w, total = hw * batch;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (c_pad == 32)
    nchw_to_nhwc_pad_kernel<32><<<grid, 256, 0, s>>>(src, c, hw, total, static_cast<__nv_bfloat16*>(dst));
  else if (c_pad == 16)
    nchw_to_nhwc_pad_kernel<16><<<grid, 256, 0, s>>>(src, c, hw, total, static_cast<__nv_bfloat16*>(dst));
  else
    return fail(VUB_E_SHAPE, "nchw_to_nhwc_pad: c_pad %d not in {16,32}", c_pad);
  return check_launch("nchw_to_nhwc_pad");
}
