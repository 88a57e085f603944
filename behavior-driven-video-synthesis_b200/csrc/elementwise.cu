// Memory-bound kernels of the VUnet path: weight-norm preparation, the K=3/30 stem, the latent
// head (sigmoid / exp / reparameterisation / KL partial sums) and the NCHW<->NHWC converters.
// All are HBM-bound: coalesced accesses, 128-bit vectors where the layout allows, warp-shuffle
// reductions.
#include <cuda_bf16.h>

#include "common.h"
#include "conv_epilogue.cuh"
#include "ptx.cuh"

namespace vub {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); every thread gets the result.
__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------------------------------------
// Weight-norm preparation (lib/modules.py:135-138,144): one block per output channel of any layer.
//   v_hat = v / ||v||_2 (bf16, packed K order (source, tap, channel)); scale = gamma*g;
//   shift = gamma*bias + beta.
// ------------------------------------------------------------------------------------------------
constexpr int kWnThreads = 128;

// One output channel `co` of one conv: normalise its weight row and emit every prepared form.  `ss_partial` >= 0:
// this thread's share of the row's sum of squares is already known (the fused Adam update computed it).
__device__ __forceinline__ void wn_emit_row(const vub_wn_entry& e, int co, float* red /*[32]*/, float ss_partial = -1.f) {
  const int len = e.cin * e.ktaps;
  const float* v = e.v + (size_t)co * len;
  float ss = 0.f;
  if (ss_partial >= 0.f) {
    ss = ss_partial;
  } else {
    for (int i = threadIdx.x; i < len; i += kWnThreads) {
      const float x = v[i];
      ss = fmaf(x, x, ss);
    }
  }
  ss = block_sum(ss, red);
  const float inv = 1.0f / sqrtf(ss);
  if (e.w_packed != nullptr) {
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(e.w_packed) + (size_t)co * e.ktot;
    for (int i = threadIdx.x; i < len; i += kWnThreads) {
      const int ci = i / e.ktaps;
      const int tap = i - ci * e.ktaps;
      const int src = ci / e.c_src;
      const int cs = ci - src * e.c_src;
      out[(src * e.ktaps + tap) * e.c_src + cs] = __float2bfloat16_rn(v[i] * inv);
    }
  }
  if (e.w_tiled != nullptr) {  // tile-major, pre-swizzled copy of the packed row (see vub_wn_entry.w_tiled)
    __nv_bfloat16* wt = static_cast<__nv_bfloat16*>(e.w_tiled);
    const int row = e.wd_col0 + co;
    for (int i = threadIdx.x; i < len; i += kWnThreads) {
      const int ci = i / e.ktaps;
      const int tap = i - ci * e.ktaps;
      const int src = ci / e.c_src;
      const int cs = ci - src * e.c_src;
      const int k = (src * e.ktaps + tap) * e.c_src + cs;
      const int kb = k >> 6, kc = k & 63;
      wt[((size_t)kb * e.tiled_rows + row) * 64 + ((((kc >> 3) ^ (row & 7)) << 3) | (kc & 7))] =
          __float2bfloat16_rn(v[i] * inv);
    }
  }
  if (e.w32 != nullptr) {  // fp32 validation path: [k][co], unscaled
    for (int i = threadIdx.x; i < len; i += kWnThreads) {
      const int ci = i / e.ktaps;
      const int tap = i - ci * e.ktaps;
      const int src = ci / e.c_src;
      const int cs = ci - src * e.c_src;
      e.w32[(size_t)((src * e.ktaps + tap) * e.c_src + cs) * e.wd_cout + e.wd_col0 + co] = v[i] * inv;
    }
  }
  if (e.wd != nullptr) {
    // data-gradient weights: row = input channel (source-major), column = (tap, output channel), scaled
    const float sc = e.gamma[co] * e.g[co] * inv;
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(e.wd);
    const size_t pitch = (size_t)e.ktaps * e.wd_cout;
    for (int i = threadIdx.x; i < len; i += kWnThreads) {
      const int ci = i / e.ktaps;
      const int tap = i - ci * e.ktaps;
      wd[(size_t)ci * pitch + (size_t)tap * e.wd_cout + e.wd_col0 + co] = __float2bfloat16_rn(v[i] * sc);
    }
    if (e.wd_tiled != nullptr) {
      __nv_bfloat16* wdt = static_cast<__nv_bfloat16*>(e.wd_tiled);
      for (int i = threadIdx.x; i < len; i += kWnThreads) {
        const int ci = i / e.ktaps;  // row
        const int tap = i - ci * e.ktaps;
        const int k = tap * e.wd_cout + e.wd_col0 + co;
        const int kb = k >> 6, kc = k & 63;
        wdt[((size_t)kb * e.cin + ci) * 64 + ((((kc >> 3) ^ (ci & 7)) << 3) | (kc & 7))] = __float2bfloat16_rn(v[i] * sc);
      }
    }
  }
  if (threadIdx.x == 0 && e.scale != nullptr) {
    const float ga = e.gamma[co];
    e.scale[co] = ga * e.g[co];
    e.shift[co] = fmaf(ga, e.bias[co], e.beta[co]);
  }
}

__global__ void __launch_bounds__(kWnThreads)
    wn_prep_kernel(const vub_wn_entry* __restrict__ table, int n_entries) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  // binary search for the entry with row_start <= row < row_start + cout
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].row_start <= row) lo = mid;
    else hi = mid - 1;
  }
  const vub_wn_entry e = table[lo];
  const int co = row - e.row_start;
  if (co >= e.cout) return;
  wn_emit_row(e, co, red);
}

// ------------------------------------------------------------------------------------------------
// Fused optimizer step: Adam on the five parameter tensors of every NormConv2d AND the
// weight-norm preparation of the updated weights for the next forward/backward, one pass over the
// parameters (the reference: torch.optim.Adam(betas=(0.5, 0.9)) over ~470 tensors,
// experiments/shape_and_pose_net.py:237-246, then the weight_norm hook of every conv in the next forward).
//   m = b1*m + (1-b1)*g;  s = b2*s + (1-b2)*g*g;  p -= (lr / (1-b1^t)) * m / (sqrt(s) / sqrt(1-b2^t) + eps)
// One block per output channel: the v row is updated, then (after a block barrier) normalised and packed
// by the same code as vub_wn_prep.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& s, float b1, float b2, float step_size,
                                             float inv_sqrt_bc2, float eps) {
  m = fmaf(b1, m, (1.0f - b1) * g);
  s = fmaf(b2, s, (1.0f - b2) * g * g);
  return p - step_size * m / (sqrtf(s) * inv_sqrt_bc2 + eps);
}

__global__ void adam_step_incr_kernel(int32_t* step) { *step += 1; }

__global__ void __launch_bounds__(kWnThreads)
    adam_wn_kernel(const vub_adam_entry* __restrict__ table, int n_entries, float b1, float b2, float eps, float bc1,
                   float inv_sqrt_bc2, const int32_t* __restrict__ step_dev) {
  __shared__ float red[32];
  if (step_dev != nullptr) {  // step count lives on the device (CUDA-graph replays): bias corrections computed here
    const float t = static_cast<float>(*step_dev);
    bc1 = 1.0f - powf(b1, t);
    inv_sqrt_bc2 = 1.0f / sqrtf(1.0f - powf(b2, t));
  }
  const int row = blockIdx.x;
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].wn.row_start <= row) lo = mid;
    else hi = mid - 1;
  }
  const vub_adam_entry e = table[lo];
  const int co = row - e.wn.row_start;
  if (co >= e.wn.cout) return;
  const float step_size = e.lr / bc1;
  const int len = e.wn.cin * e.wn.ktaps;
  const size_t off = (size_t)co * len;
  float* v = const_cast<float*>(e.wn.v) + off;
  float ss = 0.f;  // this thread's share of ||v_new||^2 (same elements it would re-read in wn_emit_row)
  for (int i = threadIdx.x; i < len; i += kWnThreads) {
    float m = e.m_v[off + i], sq = e.s_v[off + i];
    const float vn = adam_update(v[i], e.d_v[off + i], m, sq, b1, b2, step_size, inv_sqrt_bc2, eps);
    v[i] = vn;
    ss = fmaf(vn, vn, ss);
    e.m_v[off + i] = m;
    e.s_v[off + i] = sq;
  }
  if (threadIdx.x < 4) {
    // the four per-channel scalars: g, gamma, beta, bias
    float* ps[4] = {const_cast<float*>(e.wn.g), const_cast<float*>(e.wn.gamma), const_cast<float*>(e.wn.beta),
                    const_cast<float*>(e.wn.bias)};
    const float* gs[4] = {e.d_g, e.d_gamma, e.d_beta, e.d_bias};
    float* ms[4] = {e.m_g, e.m_gamma, e.m_beta, e.m_bias};
    float* ss[4] = {e.s_g, e.s_gamma, e.s_beta, e.s_bias};
    const int t = threadIdx.x;
    float m = ms[t][co], sq = ss[t][co];
    ps[t][co] = adam_update(ps[t][co], gs[t][co], m, sq, b1, b2, step_size, inv_sqrt_bc2, eps);
    ms[t][co] = m;
    ss[t][co] = sq;
  }
  __syncthreads();  // the updated row and scalars are visible to the whole block
  wn_emit_row(e.wn, co, red, ss);
}

// ------------------------------------------------------------------------------------------------
// Stem: 1x1 NormConv2d on an NCHW fp32 image -> NHWC bf16 raw + ELU (models/vunets.py:137,250).
// One thread per pixel: reads are coalesced along W per channel plane, each thread owns COUT
// accumulators and writes its pixel's COUT*2 contiguous bytes with 128-bit stores.
// ------------------------------------------------------------------------------------------------
constexpr int kStemThreads = 128;
constexpr int kStemMaxCx = 64;

// U8: the input is a uint8 HWC image batch [B, H, W, cx] (what the reference's loaders / stickman rasteriser produce
// before `ToTensor` + `x * 2 - 1`, data/__init__.py:15-43, data/data_conversions_3d.py:1154-1163); the kernel applies
// exactly that conversion — fp32 (u / 255) * 2 - 1 through a 256-entry table built with IEEE division — so the
// result is bit-identical to feeding the converted fp32 NCHW tensor, at a quarter of the input bytes and with the
// pixel's channels contiguous.
template <int COUT, bool U8>
__global__ void __launch_bounds__(kStemThreads)
    stem_kernel(const void* __restrict__ x_in, const float* __restrict__ v, const float* __restrict__ g,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ bias, int cx, long long hw, long long total_pix,
                __nv_bfloat16* __restrict__ out_raw, __nv_bfloat16* __restrict__ out_elu) {
  __shared__ float s_w[COUT * kStemMaxCx];  // [ci][co], already scaled by gamma*g/||v||
  __shared__ float s_shift[COUT];
  __shared__ float s_lut[U8 ? 256 : 1];
  const float* __restrict__ x = static_cast<const float*>(x_in);
  const uint8_t* __restrict__ x8 = static_cast<const uint8_t*>(x_in);
  if constexpr (U8) {
    for (int i = threadIdx.x; i < 256; i += kStemThreads) s_lut[i] = __fdiv_rn(static_cast<float>(i), 255.0f) * 2.0f - 1.0f;
  }
  for (int co = threadIdx.x; co < COUT; co += kStemThreads) {
    float ss = 0.f;
    for (int ci = 0; ci < cx; ++ci) ss = fmaf(v[co * cx + ci], v[co * cx + ci], ss);
    const float sc = gamma[co] * g[co] / sqrtf(ss);
    for (int ci = 0; ci < cx; ++ci) s_w[ci * COUT + co] = v[co * cx + ci] * sc;
    s_shift[co] = fmaf(gamma[co], bias[co], beta[co]);
  }
  __syncthreads();
  // Persistent: the weight-norm prologue above costs two dependent global round trips, so a block handles many
  // 128-pixel chunks (grid-stride) instead of one.
  constexpr int CPR = COUT / 8;  // 16-byte chunks per pixel
  __shared__ __align__(128) uint8_t s_out[kStemThreads * COUT * 2];
  for (long long pix0 = (long long)blockIdx.x * kStemThreads; pix0 < total_pix; pix0 += (long long)gridDim.x * kStemThreads) {
    const long long pix_raw = pix0 + threadIdx.x;
    const bool valid = pix_raw < total_pix;
    const long long pix = valid ? pix_raw : total_pix - 1;  // tail threads compute a duplicate, store nothing
    const long long n = pix / hw;
    const long long sp = pix - n * hw;
    const float* xp = x + n * cx * hw + sp;
    const uint8_t* xp8 = x8 + pix * cx;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = s_shift[co];
    for (int ci = 0; ci < cx; ++ci) {
      float xv;
      if constexpr (U8) xv = s_lut[__ldg(xp8 + ci)];
      else xv = __ldg(xp + (long long)ci * hw);
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[co] = fmaf(xv, s_w[ci * COUT + co], acc[co]);
    }
    // Coalesced stores: a thread owns one pixel (COUT*2 contiguous bytes), so direct 16-byte stores from the
    // 32 lanes of a warp hit 32 different 64/128-byte segments per instruction.  The block's 128 pixels are one
    // contiguous span of the NHWC output: stage them in shared memory (swizzled, see EpiStage) and write the span
    // linearly.
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* out = which == 0 ? out_raw : out_elu;
      if (out == nullptr) continue;
      __syncthreads();  // previous pass has been read back
      if (valid) {
#pragma unroll
        for (int j = 0; j < CPR; ++j) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = which == 0 ? acc[8 * j + i] : elu1(acc[8 * j + i]);
          __nv_bfloat162 a = __floats2bfloat162_rn(t[0], t[1]), b2 = __floats2bfloat162_rn(t[2], t[3]);
          __nv_bfloat162 c2 = __floats2bfloat162_rn(t[4], t[5]), d2 = __floats2bfloat162_rn(t[6], t[7]);
          *reinterpret_cast<uint4*>(s_out + EpiStage<COUT>::off(threadIdx.x, j)) =
              make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b2),
                         *reinterpret_cast<uint32_t*>(&c2), *reinterpret_cast<uint32_t*>(&d2));
        }
      }
      __syncthreads();
      uint4* dst = reinterpret_cast<uint4*>(out + pix0 * COUT);
#pragma unroll
      for (int k = 0; k < CPR; ++k) {
        const int c = threadIdx.x + kStemThreads * k;  // chunk index inside the block's span
        const int row = c / CPR, col = c - row * CPR;
        if (pix0 + row < total_pix) dst[c] = *reinterpret_cast<const uint4*>(s_out + EpiStage<COUT>::off(row, col));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Latent head (models/vunets.py:575,594-597; lib/losses.py:283-291).
// One thread handles 4 consecutive channels of one pixel: float4 loads of mu and logit from the
// NHWC conv output, 8-byte bf16 stores to the NHWC operands, scalar stores to the (tiny) NCHW
// outputs, warp-shuffle + block reduction of the KL summand into one partial per block.
// ------------------------------------------------------------------------------------------------
constexpr int kLatThreads = 128;

__global__ void __launch_bounds__(kLatThreads)
    latent_kernel(const float* __restrict__ ml, const float* __restrict__ eps, int c, int hw,
                  long long total4, float* __restrict__ mu_o, float* __restrict__ ls_o,
                  float* __restrict__ z_o, __nv_bfloat16* __restrict__ elu_z,
                  __nv_bfloat16* __restrict__ elu_mu, float* __restrict__ kl_partial) {
  __shared__ float red[32];
  const long long idx = (long long)blockIdx.x * kLatThreads + threadIdx.x;
  float kl = 0.f;
  if (idx < total4) {
    const int c4 = c >> 2;
    const long long pix = idx / c4;  // n*hw + sp
    const int ch = static_cast<int>(idx - pix * c4) * 4;
    const long long n = pix / hw;
    const int sp = static_cast<int>(pix - n * hw);
    const float4 m4 = *reinterpret_cast<const float4*>(ml + pix * (2 * c) + ch);
    const float4 l4 = *reinterpret_cast<const float4*>(ml + pix * (2 * c) + c + ch);
    const float mu[4] = {m4.x, m4.y, m4.z, m4.w};
    const float lg[4] = {l4.x, l4.y, l4.z, l4.w};
    float zz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long o = (n * c + ch + i) * hw + sp;  // NCHW
      const float ls = 1.0f / (1.0f + expf(-lg[i]));
      const float sd = expf(ls);
      const float e = eps != nullptr ? eps[o] : 0.f;
      const float z = fmaf(e, sd, mu[i]);
      zz[i] = z;
      mu_o[o] = mu[i];
      ls_o[o] = ls;
      z_o[o] = z;
      kl += -ls + 0.5f * (sd * sd + mu[i] * mu[i]);
    }
    if (elu_z != nullptr) {
      __nv_bfloat162 a = __floats2bfloat162_rn(elu1(zz[0]), elu1(zz[1]));
      __nv_bfloat162 b = __floats2bfloat162_rn(elu1(zz[2]), elu1(zz[3]));
      *reinterpret_cast<uint2*>(elu_z + pix * c + ch) =
          make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
    if (elu_mu != nullptr) {
      __nv_bfloat162 a = __floats2bfloat162_rn(elu1(mu[0]), elu1(mu[1]));
      __nv_bfloat162 b = __floats2bfloat162_rn(elu1(mu[2]), elu1(mu[3]));
      *reinterpret_cast<uint2*>(elu_mu + pix * c + ch) =
          make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
  }
  kl = block_sum(kl, red);
  if (threadIdx.x == 0 && kl_partial != nullptr) kl_partial[blockIdx.x] = kl;
}

// ------------------------------------------------------------------------------------------------
// Layout converters through a 32x33 shared-memory tile (coalesced on both sides).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    nchw_to_nhwc_kernel(const float* __restrict__ src, int c, int hw, __nv_bfloat16* __restrict__ raw,
                        __nv_bfloat16* __restrict__ elu) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int ch = c0 + j, px = p0 + tx;
    tile[j][tx] = (ch < c && px < hw) ? src[(n * c + ch) * hw + px] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int px = p0 + j, ch = c0 + tx;
    if (px < hw && ch < c) {
      const float v = tile[tx][j];
      const long long o = (n * hw + px) * c + ch;
      if (raw != nullptr) raw[o] = __float2bfloat16_rn(v);
      if (elu != nullptr) elu[o] = __float2bfloat16_rn(elu1(v));
    }
  }
}

__global__ void __launch_bounds__(256)
    nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int c, int hw, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int px = p0 + j, ch = c0 + tx;
    tile[j][tx] = (px < hw && ch < c) ? __bfloat162float(src[(n * hw + px) * c + ch]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int ch = c0 + j, px = p0 + tx;
    if (ch < c && px < hw) dst[(n * c + ch) * hw + px] = tile[tx][j];
  }
}

// ------------------------------------------------------------------------------------------------
// Frame epilogue of the rendering loop: scale_img(x) * 255 -> HWC uint8 (lib/utils.py:658-668,
// data/data_conversions_3d.py:1166-1173).  Same fp32 operations in the same order as the reference
// ((x + 1) / 2, clamp to [0, 1], * 255, truncation) => bit-exact.  One thread per pixel: three coalesced
// plane reads, one 3-byte store.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    frames_to_uint8_kernel(const float* __restrict__ src, long long hw, long long total, uint8_t* __restrict__ dst) {
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= total) return;
  const long long n = pix / hw, sp = pix - n * hw;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float o = (__ldg(src + (n * 3 + c) * hw + sp) + 1.0f) / 2.0f;
    o = fminf(fmaxf(o, 0.0f), 1.0f) * 255.0f;
    dst[pix * 3 + c] = static_cast<uint8_t>(static_cast<int>(o));
  }
}

}  // namespace vub

using namespace vub;

extern "C" int vub_frames_to_uint8(const float* img_nchw, int32_t batch, int32_t h, int32_t w, uint8_t* out_nhwc,
                                   void* stream) {
  clear_error();
  if (!img_nchw || !out_nhwc) return fail(VUB_E_NULL, "frames_to_uint8: NULL pointer");
  if (batch <= 0 || h <= 0 || w <= 0) return fail(VUB_E_SHAPE, "frames_to_uint8: bad shape");
  const long long hw = (long long)h * w, total = hw * batch;
  frames_to_uint8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img_nchw, hw, total, out_nhwc);
  return check_launch("frames_to_uint8");
}

extern "C" int vub_wn_prep(const vub_wn_entry* table_dev, int32_t n_entries, int32_t total_rows,
                           void* stream) {
  clear_error();
  if (!table_dev) return fail(VUB_E_NULL, "wn_prep: table is NULL");
  if (n_entries <= 0 || total_rows <= 0) return fail(VUB_E_SHAPE, "wn_prep: empty table");
  wn_prep_kernel<<<total_rows, kWnThreads, 0, static_cast<cudaStream_t>(stream)>>>(table_dev, n_entries);
  return check_launch("wn_prep");
}

extern "C" int vub_adam_wn_step(const vub_adam_entry* table_dev, int32_t n_entries, int32_t total_rows, float beta1,
                                float beta2, float eps, int32_t step, void* stream) {
  clear_error();
  if (!table_dev) return fail(VUB_E_NULL, "adam_wn_step: table is NULL");
  if (n_entries <= 0 || total_rows <= 0 || step < 1) return fail(VUB_E_SHAPE, "adam_wn_step: empty table or step < 1");
  if (!(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f)) return fail(VUB_E_SHAPE, "adam_wn_step: betas");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adam_wn_kernel<<<total_rows, kWnThreads, 0, static_cast<cudaStream_t>(stream)>>>(table_dev, n_entries, beta1, beta2, eps,
                                                                                  bc1, 1.0f / sqrtf(bc2), nullptr);
  return check_launch("adam_wn_step");
}

extern "C" int vub_adam_wn_step_dev(const vub_adam_entry* table_dev, int32_t n_entries, int32_t total_rows, float beta1,
                                    float beta2, float eps, int32_t* step_dev, void* stream) {
  clear_error();
  if (!table_dev || !step_dev) return fail(VUB_E_NULL, "adam_wn_step_dev: table or step counter is NULL");
  if (n_entries <= 0 || total_rows <= 0) return fail(VUB_E_SHAPE, "adam_wn_step_dev: empty table");
  if (!(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f)) return fail(VUB_E_SHAPE, "adam_wn_step_dev: betas");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  adam_step_incr_kernel<<<1, 1, 0, s>>>(step_dev);
  adam_wn_kernel<<<total_rows, kWnThreads, 0, s>>>(table_dev, n_entries, beta1, beta2, eps, 0.f, 0.f, step_dev);
  return check_launch("adam_wn_step_dev");
}

static int stem_fwd_launch(const void* x, bool u8, const float* v, const float* g, const float* gamma, const float* beta,
                           const float* bias, int32_t batch, int32_t cx, int32_t h, int32_t w, int32_t cout, void* out_raw,
                           void* out_elu, void* stream) {
  clear_error();
  if (!x || !v || !g || !gamma || !beta || !bias) return fail(VUB_E_NULL, "stem_fwd: NULL input");
  if (!out_raw && !out_elu) return fail(VUB_E_FLAGS, "stem_fwd: no output");
  if (batch <= 0 || h <= 0 || w <= 0 || cx <= 0 || cx > kStemMaxCx)
    return fail(VUB_E_SHAPE, "stem_fwd: batch %d cx %d h %d w %d (cx <= %d)", batch, cx, h, w, kStemMaxCx);
  if ((out_raw && !aligned16(out_raw)) || (out_elu && !aligned16(out_elu)))
    return fail(VUB_E_ALIGN, "stem_fwd: outputs must be 16B aligned");
  const long long hw = (long long)h * w, total = hw * batch;
  const long long chunks = (total + kStemThreads - 1) / kStemThreads;
  const int n_sm = sm_count();
  const long long max_grid = (long long)n_sm * 8;  // persistent blocks, each looping over 128-pixel chunks
  const unsigned grid = static_cast<unsigned>(chunks < max_grid ? chunks : max_grid);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o_raw = static_cast<__nv_bfloat16*>(out_raw);
  __nv_bfloat16* o_elu = static_cast<__nv_bfloat16*>(out_elu);
  if (cout == 32 && !u8)
    stem_kernel<32, false><<<grid, kStemThreads, 0, s>>>(x, v, g, gamma, beta, bias, cx, hw, total, o_raw, o_elu);
  else if (cout == 64 && !u8)
    stem_kernel<64, false><<<grid, kStemThreads, 0, s>>>(x, v, g, gamma, beta, bias, cx, hw, total, o_raw, o_elu);
  else if (cout == 32)
    stem_kernel<32, true><<<grid, kStemThreads, 0, s>>>(x, v, g, gamma, beta, bias, cx, hw, total, o_raw, o_elu);
  else if (cout == 64)
    stem_kernel<64, true><<<grid, kStemThreads, 0, s>>>(x, v, g, gamma, beta, bias, cx, hw, total, o_raw, o_elu);
  else
    return fail(VUB_E_SHAPE, "stem_fwd: cout %d not in {32,64}", cout);
  return check_launch("stem_fwd");
}

extern "C" int vub_stem_fwd(const float* x_nchw, const float* v, const float* g, const float* gamma,
                            const float* beta, const float* bias, int32_t batch, int32_t cx, int32_t h,
                            int32_t w, int32_t cout, void* out_raw, void* out_elu, void* stream) {
  return stem_fwd_launch(x_nchw, false, v, g, gamma, beta, bias, batch, cx, h, w, cout, out_raw, out_elu, stream);
}

extern "C" int vub_stem_fwd_u8(const uint8_t* x_nhwc_u8, const float* v, const float* g, const float* gamma,
                               const float* beta, const float* bias, int32_t batch, int32_t cx, int32_t h,
                               int32_t w, int32_t cout, void* out_raw, void* out_elu, void* stream) {
  return stem_fwd_launch(x_nhwc_u8, true, v, g, gamma, beta, bias, batch, cx, h, w, cout, out_raw, out_elu, stream);
}

extern "C" int32_t vub_latent_num_partials(int32_t batch, int32_t c, int32_t h, int32_t w) {
  const long long total4 = (long long)batch * h * w * (c / 4);
  return static_cast<int32_t>((total4 + kLatThreads - 1) / kLatThreads);
}

extern "C" int vub_latent_fwd(const float* ml_nhwc, const float* eps_nchw, int32_t batch, int32_t c,
                              int32_t h, int32_t w, float* mu_nchw, float* logstd_nchw, float* z_nchw,
                              void* elu_z_nhwc, void* elu_mu_nhwc, float* kl_partial, void* stream) {
  clear_error();
  if (!ml_nhwc || !mu_nchw || !logstd_nchw || !z_nchw) return fail(VUB_E_NULL, "latent_fwd: NULL pointer");
  if (batch <= 0 || c <= 0 || (c & 3) || h <= 0 || w <= 0)
    return fail(VUB_E_SHAPE, "latent_fwd: batch %d c %d h %d w %d (c %% 4 == 0)", batch, c, h, w);
  if (!aligned16(ml_nhwc)) return fail(VUB_E_ALIGN, "latent_fwd: ml_nhwc not 16B aligned");
  const long long total4 = (long long)batch * h * w * (c / 4);
  const unsigned grid = static_cast<unsigned>((total4 + kLatThreads - 1) / kLatThreads);
  latent_kernel<<<grid, kLatThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      ml_nhwc, eps_nchw, c, h * w, total4, mu_nchw, logstd_nchw, z_nchw,
      static_cast<__nv_bfloat16*>(elu_z_nhwc), static_cast<__nv_bfloat16*>(elu_mu_nhwc), kl_partial);
  return check_launch("latent_fwd");
}

extern "C" int vub_nchw_f32_to_nhwc_bf16(const float* src, int32_t batch, int32_t c, int32_t h,
                                         int32_t w, void* dst_raw, void* dst_elu, void* stream) {
  clear_error();
  if (!src || (!dst_raw && !dst_elu)) return fail(VUB_E_NULL, "nchw_to_nhwc: NULL pointer");
  if (batch <= 0 || c <= 0 || h <= 0 || w <= 0 || batch > 65535)
    return fail(VUB_E_SHAPE, "nchw_to_nhwc: bad shape");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, batch);
  nchw_to_nhwc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, c, hw, static_cast<__nv_bfloat16*>(dst_raw), static_cast<__nv_bfloat16*>(dst_elu));
  return check_launch("nchw_to_nhwc");
}

extern "C" int vub_nhwc_bf16_to_nchw_f32(const void* src, int32_t batch, int32_t c, int32_t h,
                                         int32_t w, float* dst, void* stream) {
  clear_error();
  if (!src || !dst) return fail(VUB_E_NULL, "nhwc_to_nchw: NULL pointer");
  if (batch <= 0 || c <= 0 || h <= 0 || w <= 0 || batch > 65535)
    return fail(VUB_E_SHAPE, "nhwc_to_nchw: bad shape");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, batch);
  nhwc_to_nchw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), c, hw, dst);
  return check_launch("nhwc_to_nchw");
}
