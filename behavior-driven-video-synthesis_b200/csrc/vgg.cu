// Memory-bound pieces of the VGG19 perceptual loss (lib/losses.py:81-119 `vgg_loss`, models/imagenet_pretrained.py:8-61
// `PerceptualVGG`): input normalisation, 2x2 max-pooling, the L1 feature distances and their gradients.  The 3x3
// convolutions + ReLU of the feature stack run on the tcgen05 conv kernels (VUB_CONV_RELU); everything here is an
// elementwise / reduction pass over NHWC bf16 feature maps: 128-bit accesses, warp-shuffle + block reductions,
// deterministic two-stage sums (per-block partials, summed by the caller).
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace vub {

constexpr int kVggThreads = 256;
__constant__ float c_vgg_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_vgg_std[3] = {0.229f, 0.224f, 0.225f};

__device__ __forceinline__ float warp_sum_v(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum_v(float v, float* red /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_v(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (warp == 0) v = warp_sum_v(v);
  return v;  // valid in warp 0
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __low2float(h[i]);
    f[2 * i + 1] = __high2float(h[i]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = pack8(f);
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// PerceptualVGG.forward's first two lines: x = (x + 1) / 2; x = (x - mean) / std, NCHW fp32 -> NHWC bf16 padded to 32
// channels (the conv kernels' smallest K block), optionally also the fp32 NCHW tensor the module returns as "input".
template <typename T>
__global__ void __launch_bounds__(kVggThreads)
    vgg_normalize_kernel(const float* __restrict__ x, long long hw, long long total, T* __restrict__ out_nhwc,
                         float* __restrict__ out_nchw) {
  const long long pix = (long long)blockIdx.x * kVggThreads + threadIdx.x;
  if (pix >= total) return;
  const long long n = pix / hw, sp = pix - n * hw;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float t = (__ldg(x + (n * 3 + c) * hw + sp) + 1.0f) / 2.0f;
    v[c] = (t - c_vgg_mean[c]) / c_vgg_std[c];
    if (out_nchw != nullptr) out_nchw[(n * 3 + c) * hw + sp] = v[c];
  }
  if (out_nhwc != nullptr) {
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8<T>(out_nhwc + pix * 32, v);
    store8<T>(out_nhwc + pix * 32 + 8, z);
    store8<T>(out_nhwc + pix * 32 + 16, z);
    store8<T>(out_nhwc + pix * 32 + 24, z);
  }
}

// L1 distance of the "input" tap and the backward of the normalisation:
//   partial[block] = sum |xn(pred) - xn(target)| = sum |p - t| / (2 std_c)
//   d_pred[n,c,h,w] = (d_xn[n,h,w,c] (gradient from conv1_1, NHWC bf16 x 32, or NULL)
//                      + g0 * scale0 * sign(xn(p) - xn(t))) / (2 std_c)
template <bool BWD, typename T>
__global__ void __launch_bounds__(kVggThreads)
    vgg_input_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long hw, long long total,
                     float* __restrict__ partial, const T* __restrict__ d_xn, const float* __restrict__ g0,
                     float scale0, float* __restrict__ d_pred) {
  __shared__ float red[32];
  const long long pix = (long long)blockIdx.x * kVggThreads + threadIdx.x;
  float acc = 0.f;
  if (pix < total) {
    const long long n = pix / hw, sp = pix - n * hw;
    float dx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (BWD) {
      if (d_xn != nullptr) load8<T>(d_xn + pix * 32, dx);
    }
    const float gs = BWD && g0 != nullptr ? __ldg(g0) * scale0 : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o = (n * 3 + c) * hw + sp;
      const float inv = 1.0f / (2.0f * c_vgg_std[c]);
      const float d = (__ldg(pred + o) - __ldg(target + o)) * inv;
      if constexpr (BWD) {
        const float sg = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.0f);
        d_pred[o] = (dx[c] + gs * sg) * inv;
      } else {
        acc += fabsf(d);
      }
    }
  }
  if constexpr (!BWD) {
    acc = block_sum_v(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
  }
}

// nn.MaxPool2d(2, 2) on NHWC bf16: one thread = 8 channels of one OUTPUT pixel.
template <typename T>
__global__ void __launch_bounds__(kVggThreads)
    maxpool2_fwd_kernel(const T* __restrict__ x, int c8, int ho, int wo, long long total, T* __restrict__ y) {
  const long long idx = (long long)blockIdx.x * kVggThreads + threadIdx.x;
  if (idx >= total) return;
  const long long opix = idx / c8;
  const int ch = static_cast<int>(idx - opix * c8) * 8;
  const int ox = static_cast<int>(opix % wo);
  const long long t = opix / wo;
  const int oy = static_cast<int>(t % ho);
  const long long n = t / ho;
  const int C = c8 * 8, W = 2 * wo;
  const T* p = x + ((n * (2 * ho) + 2 * oy) * W + 2 * ox) * C + ch;
  float a[8], b[8], m[8];
  load8<T>(p, m);
  load8<T>(p + C, a);
  load8<T>(p + (long long)W * C, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(fmaxf(m[i], a[i]), b[i]);
  load8<T>(p + (long long)W * C + C, a);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], a[i]);
  store8<T>(y + opix * C + ch, m);
}

// Backward of max-pooling FUSED with the ReLU in front of it and the tap gradient:
//   dx[n,h,w,c] = (x is the FIRST maximum of its window in scan order ? dy[n,h/2,w/2,c] : 0) * (x > 0) + addend
// (x = the ReLU output that was pooled, so dx is the gradient w.r.t. the conv's pre-activation; autograd's max_pool2d
// backward also routes to the first maximum).  One thread = 8 channels of one window.
template <typename T>
__global__ void __launch_bounds__(kVggThreads)
    maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ addend, int c8, int ho,
                        int wo, long long total, T* __restrict__ dx) {
  const long long idx = (long long)blockIdx.x * kVggThreads + threadIdx.x;
  if (idx >= total) return;
  const long long opix = idx / c8;
  const int ch = static_cast<int>(idx - opix * c8) * 8;
  const int ox = static_cast<int>(opix % wo);
  const long long t = opix / wo;
  const int oy = static_cast<int>(t % ho);
  const long long n = t / ho;
  const int C = c8 * 8, W = 2 * wo;
  const long long base = ((n * (2 * ho) + 2 * oy) * W + 2 * ox) * C + ch;
  const long long offs[4] = {0, C, (long long)W * C, (long long)W * C + C};
  float v[4][8], g[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) load8<T>(x + base + offs[k], v[k]);
  load8<T>(dy + opix * C + ch, g);
  int arg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    arg[i] = 0;
    float m = v[0][i];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (v[k][i] > m) {
        m = v[k][i];
        arg[i] = k;
      }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float o[8];
    if (addend != nullptr) load8<T>(addend + base + offs[k], o);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] += (arg[i] == k && v[k][i] > 0.f) ? g[i] : 0.f;
    store8<T>(dx + base + offs[k], o);
  }
}

// torch.mean(torch.abs(tf - pf)) of a feature tap: per-block partial sums of |pf - tf| (8 elements per thread).
template <typename T>
__global__ void __launch_bounds__(kVggThreads)
    l1_partial_kernel(const T* __restrict__ pf, const T* __restrict__ tf, long long n8, float* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kVggThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kVggThreads) {
    float a[8], b[8];
    load8<T>(pf + 8 * i, a);
    load8<T>(tf + 8 * i, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += fabsf(a[k] - b[k]);
  }
  acc = block_sum_v(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Gradient of that term w.r.t. the conv's PRE-activation (pf = ReLU(u)): g * scale * sign(pf - tf) * (pf > 0),
// g = the upstream gradient of the loss term (a device scalar: no host synchronisation), scale = weight / numel.
template <typename T>
__global__ void __launch_bounds__(kVggThreads)
    l1_grad_kernel(const T* __restrict__ pf, const T* __restrict__ tf, long long n8, const float* __restrict__ g,
                   float scale, T* __restrict__ out) {
  const long long i = (long long)blockIdx.x * kVggThreads + threadIdx.x;
  if (i >= n8) return;
  const float gs = __ldg(g) * scale;
  float a[8], b[8], o[8];
  load8<T>(pf + 8 * i, a);
  load8<T>(tf + 8 * i, b);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float d = a[k] - b[k];
    o[k] = a[k] > 0.f ? (d > 0.f ? gs : (d < 0.f ? -gs : 0.f)) : 0.f;
  }
  store8<T>(out + 8 * i, o);
}

}  // namespace vub

using namespace vub;

static unsigned blocks_for(long long n) { return static_cast<unsigned>((n + kVggThreads - 1) / kVggThreads); }

extern "C" int vub_vgg_normalize(const float* x_nchw, int32_t batch, int32_t h, int32_t w, void* out_nhwc32,
                                 float* out_nchw, int32_t f32_acts, void* stream) {
  clear_error();
  if (!x_nchw || (!out_nhwc32 && !out_nchw)) return fail(VUB_E_NULL, "vgg_normalize: NULL pointer");
  if (batch <= 0 || h <= 0 || w <= 0) return fail(VUB_E_SHAPE, "vgg_normalize: bad shape");
  if (out_nhwc32 && !aligned16(out_nhwc32)) return fail(VUB_E_ALIGN, "vgg_normalize: output not 16B aligned");
  const long long hw = (long long)h * w, total = hw * batch;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    vgg_normalize_kernel<float><<<blocks_for(total), kVggThreads, 0, st>>>(x_nchw, hw, total,
                                                                           static_cast<float*>(out_nhwc32), out_nchw);
  else
    vgg_normalize_kernel<__nv_bfloat16><<<blocks_for(total), kVggThreads, 0, st>>>(
        x_nchw, hw, total, static_cast<__nv_bfloat16*>(out_nhwc32), out_nchw);
  return check_launch("vgg_normalize");
}

extern "C" int32_t vub_vgg_input_l1_partials(int32_t batch, int32_t h, int32_t w) {
  return static_cast<int32_t>(blocks_for((long long)batch * h * w));
}

extern "C" int vub_vgg_input_l1(const float* pred_nchw, const float* target_nchw, int32_t batch, int32_t h, int32_t w,
                                float* partial, void* stream) {
  clear_error();
  if (!pred_nchw || !target_nchw || !partial) return fail(VUB_E_NULL, "vgg_input_l1: NULL pointer");
  if (batch <= 0 || h <= 0 || w <= 0) return fail(VUB_E_SHAPE, "vgg_input_l1: bad shape");
  const long long hw = (long long)h * w, total = hw * batch;
  vgg_input_kernel<false, float><<<blocks_for(total), kVggThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pred_nchw, target_nchw, hw, total, partial, nullptr, nullptr, 0.f, nullptr);
  return check_launch("vgg_input_l1");
}

extern "C" int vub_vgg_input_bwd(const float* pred_nchw, const float* target_nchw, const void* d_xn_nhwc32,
                                 const float* g0_dev, float scale0, int32_t batch, int32_t h, int32_t w,
                                 float* d_pred_nchw, int32_t f32_acts, void* stream) {
  clear_error();
  if (!pred_nchw || !target_nchw || !d_pred_nchw) return fail(VUB_E_NULL, "vgg_input_bwd: NULL pointer");
  if (batch <= 0 || h <= 0 || w <= 0) return fail(VUB_E_SHAPE, "vgg_input_bwd: bad shape");
  if (d_xn_nhwc32 && !aligned16(d_xn_nhwc32)) return fail(VUB_E_ALIGN, "vgg_input_bwd: d_xn not 16B aligned");
  const long long hw = (long long)h * w, total = hw * batch;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    vgg_input_kernel<true, float><<<blocks_for(total), kVggThreads, 0, st>>>(
        pred_nchw, target_nchw, hw, total, nullptr, static_cast<const float*>(d_xn_nhwc32), g0_dev, scale0, d_pred_nchw);
  else
    vgg_input_kernel<true, __nv_bfloat16><<<blocks_for(total), kVggThreads, 0, st>>>(
        pred_nchw, target_nchw, hw, total, nullptr, static_cast<const __nv_bfloat16*>(d_xn_nhwc32), g0_dev, scale0,
        d_pred_nchw);
  return check_launch("vgg_input_bwd");
}

extern "C" int vub_maxpool2_fwd(const void* x_nhwc, int32_t batch, int32_t h, int32_t w, int32_t c, void* y_nhwc,
                                int32_t f32_acts, void* stream) {
  clear_error();
  if (!x_nhwc || !y_nhwc) return fail(VUB_E_NULL, "maxpool2_fwd: NULL pointer");
  if (batch <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || c % 8)
    return fail(VUB_E_SHAPE, "maxpool2_fwd: batch %d h %d w %d c %d (h, w even; c %% 8 == 0)", batch, h, w, c);
  if (!aligned16(x_nhwc) || !aligned16(y_nhwc)) return fail(VUB_E_ALIGN, "maxpool2_fwd: pointers must be 16B aligned");
  const long long total = (long long)batch * (h / 2) * (w / 2) * (c / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    maxpool2_fwd_kernel<float><<<blocks_for(total), kVggThreads, 0, st>>>(static_cast<const float*>(x_nhwc), c / 8, h / 2,
                                                                          w / 2, total, static_cast<float*>(y_nhwc));
  else
    maxpool2_fwd_kernel<__nv_bfloat16><<<blocks_for(total), kVggThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x_nhwc), c / 8, h / 2, w / 2, total, static_cast<__nv_bfloat16*>(y_nhwc));
  return check_launch("maxpool2_fwd");
}

extern "C" int vub_maxpool2_relu_bwd(const void* x_nhwc, const void* dy_nhwc, const void* addend_nhwc, int32_t batch,
                                     int32_t h, int32_t w, int32_t c, void* dx_nhwc, int32_t f32_acts, void* stream) {
  clear_error();
  if (!x_nhwc || !dy_nhwc || !dx_nhwc) return fail(VUB_E_NULL, "maxpool2_relu_bwd: NULL pointer");
  if (batch <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || c % 8)
    return fail(VUB_E_SHAPE, "maxpool2_relu_bwd: batch %d h %d w %d c %d (h, w even; c %% 8 == 0)", batch, h, w, c);
  const void* ptrs[] = {x_nhwc, dy_nhwc, addend_nhwc, dx_nhwc};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(VUB_E_ALIGN, "maxpool2_relu_bwd: pointer %p not 16B aligned", q);
  const long long total = (long long)batch * (h / 2) * (w / 2) * (c / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    maxpool2_bwd_kernel<float><<<blocks_for(total), kVggThreads, 0, st>>>(
        static_cast<const float*>(x_nhwc), static_cast<const float*>(dy_nhwc), static_cast<const float*>(addend_nhwc),
        c / 8, h / 2, w / 2, total, static_cast<float*>(dx_nhwc));
  else
    maxpool2_bwd_kernel<__nv_bfloat16><<<blocks_for(total), kVggThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x_nhwc), static_cast<const __nv_bfloat16*>(dy_nhwc),
        static_cast<const __nv_bfloat16*>(addend_nhwc), c / 8, h / 2, w / 2, total, static_cast<__nv_bfloat16*>(dx_nhwc));
  return check_launch("maxpool2_relu_bwd");
}

extern "C" int32_t vub_l1_num_partials(int64_t numel) {
  const long long n8 = numel / 8;
  const long long want = (n8 + kVggThreads - 1) / kVggThreads;
  const long long cap = 4LL * sm_count();
  return static_cast<int32_t>(want < 1 ? 1 : (want < cap ? want : cap));
}

extern "C" int vub_l1_partial(const void* pf, const void* tf, int64_t numel, float* partial, int32_t f32_acts,
                              void* stream) {
  clear_error();
  if (!pf || !tf || !partial) return fail(VUB_E_NULL, "l1_partial: NULL pointer");
  if (numel <= 0 || numel % 8) return fail(VUB_E_SHAPE, "l1_partial: numel %lld (multiple of 8)", (long long)numel);
  if (!aligned16(pf) || !aligned16(tf)) return fail(VUB_E_ALIGN, "l1_partial: pointers must be 16B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    l1_partial_kernel<float><<<vub_l1_num_partials(numel), kVggThreads, 0, st>>>(
        static_cast<const float*>(pf), static_cast<const float*>(tf), numel / 8, partial);
  else
    l1_partial_kernel<__nv_bfloat16><<<vub_l1_num_partials(numel), kVggThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(pf), static_cast<const __nv_bfloat16*>(tf), numel / 8, partial);
  return check_launch("l1_partial");
}

extern "C" int vub_l1_relu_grad(const void* pf, const void* tf, int64_t numel, const float* g_dev, float scale, void* out,
                                int32_t f32_acts, void* stream) {
  clear_error();
  if (!pf || !tf || !g_dev || !out) return fail(VUB_E_NULL, "l1_relu_grad: NULL pointer");
  if (numel <= 0 || numel % 8) return fail(VUB_E_SHAPE, "l1_relu_grad: numel %lld (multiple of 8)", (long long)numel);
  if (!aligned16(pf) || !aligned16(tf) || !aligned16(out))
    return fail(VUB_E_ALIGN, "l1_relu_grad: pointers must be 16B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32_acts)
    l1_grad_kernel<float><<<blocks_for(numel / 8), kVggThreads, 0, st>>>(
        static_cast<const float*>(pf), static_cast<const float*>(tf), numel / 8, g_dev, scale, static_cast<float*>(out));
  else
    l1_grad_kernel<__nv_bfloat16><<<blocks_for(numel / 8), kVggThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(pf), static_cast<const __nv_bfloat16*>(tf), numel / 8, g_dev, scale,
        static_cast<__nv_bfloat16*>(out));
  return check_launch("l1_relu_grad");
}
