// Latent bookkeeping of the VunetOrg family (models/vunets.py:151-156 latent_sample, 704-760 the 4-group
// autoregressive prior with its SpaceToDepth / DepthToSpace regrouping, lib/modules.py:11-34).
//
// In NHWC the reference's channel regrouping is pixel regrouping: group g = (gi, gj) of SpaceToDepth(x) is the
// parity sub-grid x[:, gi::2, gj::2, :], and DepthToSpace(cat(groups)) puts group g back on that sub-grid.  So the
// split / merge of the prior groups costs no pass of its own: the sampling kernel scatters every group straight to
// its parity positions of the merged maps, and the regroup kernel gathers parity sub-grids (optionally through ELU).
// Memory-bound elementwise work on the two smallest latent maps (<= 128 channels x 16^2 per sample).
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace vub {

constexpr int kLsThreads = 128;

template <typename T>
__device__ __forceinline__ void store4(T* dst, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* dst, const float (&v)[4]) {
  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
template <typename T>
__device__ __forceinline__ void load4(const T* src, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* src, float (&v)[4]) {
  const float4 f = *reinterpret_cast<const float4*>(src);
  v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* src, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(src);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  v[0] = __low2float(a), v[1] = __high2float(a), v[2] = __low2float(b), v[3] = __high2float(b);
}

// One thread = 4 consecutive channels of one pixel of the (h x w) parameter map.
template <typename ACT>
__global__ void __launch_bounds__(kLsThreads)
    latent_sample_kernel(const float* __restrict__ p, const float* __restrict__ eps, int c, int h, int w, int sub, int gi,
                         int gj, long long total4, float* __restrict__ mean_o, float* __restrict__ z_o,
                         ACT* __restrict__ z_raw, ACT* __restrict__ z_elu, ACT* __restrict__ grp_elu) {
  const long long idx = (long long)blockIdx.x * kLsThreads + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = c >> 2;
  const long long pix = idx / c4;  // (n*h + y)*w + x
  const int ch = static_cast<int>(idx - pix * c4) * 4;
  const int x = static_cast<int>(pix % w);
  const long long t = pix / w;
  const int y = static_cast<int>(t % h);
  const long long n = t / h;
  float pv[4], zv[4], ev[4];
  load4<float>(p + pix * c + ch, pv);
  const int H = h * sub, W = w * sub, Y = y * sub + gi, X = x * sub + gj;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float e = eps != nullptr ? eps[((n * c + ch + i) * h + y) * w + x] : 0.f;
    zv[i] = pv[i] + e;  // latent_sample: mean + 1.0 * eps
    ev[i] = elu1(zv[i]);
    const long long o = ((n * c + ch + i) * H + Y) * W + X;  // NCHW, merged map
    if (mean_o != nullptr) mean_o[o] = pv[i];
    if (z_o != nullptr) z_o[o] = zv[i];
  }
  const long long mo = ((n * H + Y) * W + X) * c + ch;  // NHWC, merged map
  if (z_raw != nullptr) store4<ACT>(z_raw + mo, zv);
  if (z_elu != nullptr) store4<ACT>(z_elu + mo, ev);
  if (grp_elu != nullptr) store4<ACT>(grp_elu + pix * c + ch, ev);
}

// Backward of latent_sample: d_p[n, y, x, ch] = d_mean[n, ch, Y, X] (NCHW fp32, merged map) + dz[n, Y, X, ch] (NHWC
// activation gradient, merged map), (Y, X) = (y*sub + gi, x*sub + gj).  Either input may be absent.
template <typename ACT>
__global__ void __launch_bounds__(kLsThreads)
    latent_sample_bwd_kernel(const float* __restrict__ d_mean, const ACT* __restrict__ dz, int c, int h, int w, int sub,
                             int gi, int gj, long long total4, ACT* __restrict__ d_p) {
  const long long idx = (long long)blockIdx.x * kLsThreads + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = c >> 2;
  const long long pix = idx / c4;
  const int ch = static_cast<int>(idx - pix * c4) * 4;
  const int x = static_cast<int>(pix % w);
  const long long t = pix / w;
  const int y = static_cast<int>(t % h);
  const long long n = t / h;
  const int H = h * sub, W = w * sub, Y = y * sub + gi, X = x * sub + gj;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (dz != nullptr) load4<ACT>(dz + ((n * H + Y) * W + X) * c + ch, v);
  if (d_mean != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += d_mean[((n * c + ch + i) * H + Y) * W + X];
  }
  store4<ACT>(d_p + pix * c + ch, v);
}

// SpaceToDepth(2) on an NHWC activation: one thread = 4 channels of one INPUT pixel.
//   group_major == 0: out[n, y/2, x/2, ((y%2)*2 + x%2)*c + ch]      (the reference's channel order)
//   group_major == 1: out[(y%2)*2 + x%2][n, y/2, x/2, ch]           (torch.split of the above, each group compact)
// inverse != 0 runs the same index map backwards (DepthToSpace / merge of the groups); accumulate != 0 adds to
// dst_copy instead of overwriting it (training: the regrouping's backward is the inverse regrouping of the gradient).
template <typename ACT>
__global__ void __launch_bounds__(kLsThreads)
    regroup_kernel(const ACT* __restrict__ src, int c, int h, int w, long long total4, long long group_stride,
                   int group_major, int inverse, int want_elu, int accumulate, ACT* __restrict__ dst_copy,
                   ACT* __restrict__ dst_elu) {
  const long long idx = (long long)blockIdx.x * kLsThreads + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = c >> 2;
  const long long pix = idx / c4;
  const int ch = static_cast<int>(idx - pix * c4) * 4;
  const int x = static_cast<int>(pix % w);
  const long long t = pix / w;
  const int y = static_cast<int>(t % h);
  const long long n = t / h;
  const int g = (y & 1) * 2 + (x & 1);
  const long long half = (n * (h >> 1) + (y >> 1)) * (w >> 1) + (x >> 1);
  const long long full = pix * c + ch;
  const long long grouped = group_major ? g * group_stride + half * c + ch : (half * 4 + g) * c + ch;
  float v[4];
  load4<ACT>(src + (inverse ? grouped : full), v);
  const long long o = inverse ? full : grouped;
  if (dst_copy != nullptr) {
    if (accumulate) {  // backward of the regrouping: add to a gradient that already has other contributions
      float old[4];
      load4<ACT>(dst_copy + o, old);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += old[i];
    }
    store4<ACT>(dst_copy + o, v);
  }
  if (want_elu) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = elu1(v[i]);
    store4<ACT>(dst_elu + o, v);
  }
}

}  // namespace vub

using namespace vub;

extern "C" int vub_latent_sample(const float* p_nhwc, const float* eps_nchw, int32_t batch, int32_t c, int32_t h, int32_t w,
                                 int32_t sub, int32_t gi, int32_t gj, float* mean_nchw, float* z_nchw, void* z_raw_nhwc,
                                 void* z_elu_nhwc, void* grp_elu_nhwc, int32_t f32_acts, void* stream) {
  clear_error();
  if (!p_nhwc) return fail(VUB_E_NULL, "latent_sample: p is NULL");
  if (batch <= 0 || c <= 0 || c % 4 || h <= 0 || w <= 0)
    return fail(VUB_E_SHAPE, "latent_sample: batch %d c %d h %d w %d (c %% 4 == 0)", batch, c, h, w);
  if ((sub != 1 && sub != 2) || gi < 0 || gj < 0 || gi >= sub || gj >= sub)
    return fail(VUB_E_SHAPE, "latent_sample: sub %d group (%d, %d)", sub, gi, gj);
  const void* ptrs[] = {p_nhwc, z_raw_nhwc, z_elu_nhwc, grp_elu_nhwc};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(VUB_E_ALIGN, "latent_sample: pointer %p not 16B aligned", q);
  const long long total4 = (long long)batch * h * w * (c / 4);
  const unsigned blocks = static_cast<unsigned>((total4 + kLsThreads - 1) / kLsThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    latent_sample_kernel<float><<<blocks, kLsThreads, 0, st>>>(p_nhwc, eps_nchw, c, h, w, sub, gi, gj, total4, mean_nchw,
                                                               z_nchw, static_cast<float*>(z_raw_nhwc),
                                                               static_cast<float*>(z_elu_nhwc),
                                                               static_cast<float*>(grp_elu_nhwc));
  else
    latent_sample_kernel<__nv_bfloat16><<<blocks, kLsThreads, 0, st>>>(
        p_nhwc, eps_nchw, c, h, w, sub, gi, gj, total4, mean_nchw, z_nchw, static_cast<__nv_bfloat16*>(z_raw_nhwc),
        static_cast<__nv_bfloat16*>(z_elu_nhwc), static_cast<__nv_bfloat16*>(grp_elu_nhwc));
  return check_launch("latent_sample");
}

extern "C" int vub_space_to_depth(const void* src, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t group_major,
                                  int32_t inverse, int32_t accumulate, void* dst_copy, void* dst_elu, int32_t f32_acts,
                                  void* stream) {
  clear_error();
  if (!src || (!dst_copy && !dst_elu)) return fail(VUB_E_NULL, "space_to_depth: NULL pointer");
  if (accumulate && (!dst_copy || dst_elu)) return fail(VUB_E_FLAGS, "space_to_depth: accumulate needs dst_copy only");
  if (batch <= 0 || c <= 0 || c % 8 || h <= 0 || w <= 0 || (h & 1) || (w & 1))
    return fail(VUB_E_SHAPE, "space_to_depth: batch %d h %d w %d c %d (h, w even; c %% 8 == 0)", batch, h, w, c);
  const void* ptrs[] = {src, dst_copy, dst_elu};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(VUB_E_ALIGN, "space_to_depth: pointer %p not 16B aligned", q);
  const long long total4 = (long long)batch * h * w * (c / 4);
  const long long group_stride = (long long)batch * (h / 2) * (w / 2) * c;
  const unsigned blocks = static_cast<unsigned>((total4 + kLsThreads - 1) / kLsThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    regroup_kernel<float><<<blocks, kLsThreads, 0, st>>>(static_cast<const float*>(src), c, h, w, total4, group_stride,
                                                         group_major, inverse, dst_elu != nullptr, accumulate,
                                                         static_cast<float*>(dst_copy), static_cast<float*>(dst_elu));
  else
    regroup_kernel<__nv_bfloat16><<<blocks, kLsThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(src), c, h, w, total4, group_stride, group_major, inverse, dst_elu != nullptr,
        accumulate, static_cast<__nv_bfloat16*>(dst_copy), static_cast<__nv_bfloat16*>(dst_elu));
  return check_launch("space_to_depth");
}

extern "C" int vub_latent_sample_bwd(const float* d_mean_nchw, const void* dz_nhwc, int32_t batch, int32_t c, int32_t h,
                                     int32_t w, int32_t sub, int32_t gi, int32_t gj, void* d_p_nhwc, int32_t f32_acts,
                                     void* stream) {
  clear_error();
  if (!d_p_nhwc || (!d_mean_nchw && !dz_nhwc)) return fail(VUB_E_NULL, "latent_sample_bwd: NULL pointer");
  if (batch <= 0 || c <= 0 || c % 4 || h <= 0 || w <= 0)
    return fail(VUB_E_SHAPE, "latent_sample_bwd: batch %d c %d h %d w %d (c %% 4 == 0)", batch, c, h, w);
  if ((sub != 1 && sub != 2) || gi < 0 || gj < 0 || gi >= sub || gj >= sub)
    return fail(VUB_E_SHAPE, "latent_sample_bwd: sub %d group (%d, %d)", sub, gi, gj);
  if (!aligned16(d_p_nhwc) || (dz_nhwc && !aligned16(dz_nhwc)))
    return fail(VUB_E_ALIGN, "latent_sample_bwd: pointers must be 16B aligned");
  const long long total4 = (long long)batch * h * w * (c / 4);
  const unsigned blocks = static_cast<unsigned>((total4 + kLsThreads - 1) / kLsThreads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    latent_sample_bwd_kernel<float><<<blocks, kLsThreads, 0, st>>>(d_mean_nchw, static_cast<const float*>(dz_nhwc), c, h, w,
                                                                   sub, gi, gj, total4, static_cast<float*>(d_p_nhwc));
  else
    latent_sample_bwd_kernel<__nv_bfloat16><<<blocks, kLsThreads, 0, st>>>(
        d_mean_nchw, static_cast<const __nv_bfloat16*>(dz_nhwc), c, h, w, sub, gi, gj, total4,
        static_cast<__nv_bfloat16*>(d_p_nhwc));
  return check_launch("latent_sample_bwd");
}
