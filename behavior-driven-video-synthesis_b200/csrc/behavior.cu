// Elementwise pieces of behavior-net sampling (BASELINE.json configs[3]): the reverse pass of the latent flow
// (models/flow/blocks.py:531-559 UnconditionalFlatDoubleCouplingFlowBlock2: Shuffle 692-704, DoubleVectorCouplingBlock2
// 276-319, ActNorm lib/modules.py:260-330) and the residual LSTM decoder of models/pose_behavior_rnn.py:463-506, 603-626
// (nn.LSTMCell + Linear, x_{t+1} = x_t + n_out(h_t)).  All matrix products (the coupling MLPs, the LSTM gates, n_out) run
// as 1x1 "convolutions" on [batch, 1, 1, features] tensors through vub_conv_fwd (tcgen05, split-K clusters); what is
// here touches [batch, <= 4096] fp32 rows: latency-bound glue, one thread per element (vectorised by 4 where aligned).
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace vub {

constexpr int kBhThreads = 256;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// y[b, j] = x[b, idx[j]] * a[idx[j]] + off[idx[j]]   (idx NULL = identity, a NULL = 1, off NULL = 0)
// = Shuffle.reverse (x[:, backward_shuffle_idx]) fused with the ActNorm.reverse (h / scale - loc) of the block before
// it: a = 1 / scale, off = -loc.  Writes fp32 and, for columns [bf_col0, bf_col0 + bf_cols), a compact bf16 copy
// (the half the next coupling MLPs read).
__global__ void __launch_bounds__(kBhThreads)
    gather_affine_kernel(const float* __restrict__ x, const int* __restrict__ idx, const float* __restrict__ scale,
                         const float* __restrict__ loc, int d, long long total, float* __restrict__ y,
                         __nv_bfloat16* __restrict__ y_bf, int bf_col0, int bf_cols) {
  const long long i = (long long)blockIdx.x * kBhThreads + threadIdx.x;
  if (i >= total) return;
  const long long b = i / d;
  const int j = static_cast<int>(i - b * d);
  const int s = idx != nullptr ? idx[j] : j;
  float v = x[b * d + s];
  if (scale != nullptr) v = v / scale[s] - loc[s];
  y[i] = v;
  if (y_bf != nullptr && j >= bf_col0 && j < bf_col0 + bf_cols) y_bf[b * bf_cols + (j - bf_col0)] = __float2bfloat16_rn(v);
}

// Reverse affine coupling: out = (x_keep - t) * exp(-tanh(s_raw))  (the s-net ends in nn.Tanh: use_tanh=True), written
// into columns [col0, col0 + d2) of the fp32 state row (pitch d) and as a compact bf16 copy.
__global__ void __launch_bounds__(kBhThreads)
    couple_inv_kernel(const float* __restrict__ state_in, const float* __restrict__ s_raw, const float* __restrict__ t,
                      int d, int d2, int col0, long long total, float* __restrict__ state_out,
                      __nv_bfloat16* __restrict__ out_bf) {
  const long long i = (long long)blockIdx.x * kBhThreads + threadIdx.x;
  if (i >= total) return;
  const long long b = i / d2;
  const int j = static_cast<int>(i - b * d2);
  const float xk = state_in[b * d + col0 + j];
  const float v = (xk - t[i]) * expf(-tanhf(s_raw[i]));
  state_out[b * d + col0 + j] = v;
  if (out_bf != nullptr) out_bf[i] = __float2bfloat16_rn(v);
}

// nn.LSTMCell update from the pre-activations gates[b, 4H] = W_ih x + b_ih + W_hh h + b_hh (chunks i, f, g, o):
//   c' = sigmoid(f) * c + sigmoid(i) * tanh(g);  h' = sigmoid(o) * tanh(c')
// c is updated in place (fp32); h' is written as bf16 twice: compact [B, H] (input of n_out) and into columns [0, H) of
// the next step's concatenated gate input [B, H + xpad].
__global__ void __launch_bounds__(kBhThreads)
    lstm_cell_kernel(const float* __restrict__ gates, float* __restrict__ c, int hdim, long long total,
                     __nv_bfloat16* __restrict__ h_lin, __nv_bfloat16* __restrict__ h_cat, int cat_pitch,
                     float* __restrict__ h_f32) {
  const long long i = (long long)blockIdx.x * kBhThreads + threadIdx.x;
  if (i >= total) return;
  const long long b = i / hdim;
  const int j = static_cast<int>(i - b * hdim);
  const float* g = gates + b * 4 * hdim;
  const float ig = sigmoidf_(g[j]), fg = sigmoidf_(g[hdim + j]), gg = tanhf(g[2 * hdim + j]), og = sigmoidf_(g[3 * hdim + j]);
  const float cn = fg * c[i] + ig * gg;
  c[i] = cn;
  const float hf = og * tanhf(cn);
  const __nv_bfloat16 h = __float2bfloat16_rn(hf);
  if (h_lin != nullptr) h_lin[i] = h;
  h_cat[b * cat_pitch + j] = h;
  if (h_f32 != nullptr) h_f32[i] = hf;
}

// BEncoder's nn.LSTM over a given sequence (models/pose_behavior_rnn.py:177-178): the step's input x[:, t] (fp32, row
// pitch x_pitch floats) as bf16, zero-padded to xpad, into columns [hdim, hdim + xpad) of the gate input.
__global__ void __launch_bounds__(kBhThreads)
    lstm_input_kernel(const float* __restrict__ x_t, long long x_pitch, int n, int xpad, long long total,
                      __nv_bfloat16* __restrict__ cat, int cat_pitch, int hdim) {
  const long long i = (long long)blockIdx.x * kBhThreads + threadIdx.x;
  if (i >= total) return;
  const long long b = i / xpad;
  const int j = static_cast<int>(i - b * xpad);
  cat[b * cat_pitch + hdim + j] = __float2bfloat16_rn(j < n ? x_t[b * x_pitch + j] : 0.f);
}

// ResidualRNNDecoder.forward's `out + res, res` for step t: x_{t+1} = x_t + v (v = n_out(h), fp32 [B, vpitch]):
//   xs[b, t, :] = x_{t+1}, cs[b, t, :] = x_t, x (state, fp32 [B, n]) <- x_{t+1}, and the bf16 copy of x_{t+1} into columns
//   [hdim, hdim + xpad) of the next gate input (zero beyond n).
__global__ void __launch_bounds__(kBhThreads)
    decoder_step_kernel(float* __restrict__ x, const float* __restrict__ v, int n, int vpitch, int xpad, int t, int steps,
                        long long total, float* __restrict__ xs, float* __restrict__ cs, __nv_bfloat16* __restrict__ x_cat,
                        int cat_pitch, int hdim) {
  const long long i = (long long)blockIdx.x * kBhThreads + threadIdx.x;
  if (i >= total) return;
  const long long b = i / xpad;
  const int j = static_cast<int>(i - b * xpad);
  float xn = 0.f;
  if (j < n) {
    const float xo = x[b * n + j];
    xn = xo + v[b * vpitch + j];
    x[b * n + j] = xn;
    xs[(b * steps + t) * n + j] = xn;
    cs[(b * steps + t) * n + j] = xo;
  }
  x_cat[b * cat_pitch + hdim + j] = __float2bfloat16_rn(xn);
}

// Start of a sequence: hidden = cell = b (generate_seq: set_hidden(hidden=b, cell=b)), x = x_start.
__global__ void __launch_bounds__(kBhThreads)
    decoder_init_kernel(const float* __restrict__ b_lat, const float* __restrict__ x0, int hdim, int n, int xpad,
                        long long batch, float* __restrict__ c, float* __restrict__ x, __nv_bfloat16* __restrict__ cat,
                        int cat_pitch) {
  const long long i = (long long)blockIdx.x * kBhThreads + threadIdx.x;
  const long long total = batch * cat_pitch;
  if (i >= total) return;
  const long long b = i / cat_pitch;
  const int j = static_cast<int>(i - b * cat_pitch);
  if (j < hdim) {
    const float v = b_lat[b * hdim + j];
    c[b * hdim + j] = v;
    cat[i] = __float2bfloat16_rn(v);
  } else {
    const int k = j - hdim;
    float v = 0.f;
    if (k < n) {
      v = x0[b * n + k];
      x[b * n + k] = v;
    }
    cat[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace vub

using namespace vub;

static unsigned bh_blocks(long long n) { return static_cast<unsigned>((n + kBhThreads - 1) / kBhThreads); }

extern "C" int vub_gather_affine(const float* x, const int32_t* idx, const float* scale, const float* loc, int32_t batch,
                                 int32_t d, float* y, void* y_bf16, int32_t bf_col0, int32_t bf_cols, void* stream) {
  clear_error();
  if (!x || !y || (scale != nullptr) != (loc != nullptr)) return fail(VUB_E_NULL, "gather_affine: NULL pointer / scale without loc");
  if (batch <= 0 || d <= 0 || (y_bf16 && (bf_col0 < 0 || bf_cols <= 0 || bf_col0 + bf_cols > d)))
    return fail(VUB_E_SHAPE, "gather_affine: batch %d d %d bf [%d, +%d)", batch, d, bf_col0, bf_cols);
  if (x == y) return fail(VUB_E_FLAGS, "gather_affine: not in place");
  const long long total = (long long)batch * d;
  gather_affine_kernel<<<bh_blocks(total), kBhThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, idx, scale, loc, d, total, y, static_cast<__nv_bfloat16*>(y_bf16), bf_col0, bf_cols);
  return check_launch("gather_affine");
}

extern "C" int vub_couple_inv(const float* state_in, const float* s_raw, const float* t, int32_t batch, int32_t d,
                              int32_t d2, int32_t col0, float* state_out, void* out_bf16, void* stream) {
  clear_error();
  if (!state_in || !s_raw || !t || !state_out) return fail(VUB_E_NULL, "couple_inv: NULL pointer");
  if (batch <= 0 || d <= 0 || d2 <= 0 || col0 < 0 || col0 + d2 > d) return fail(VUB_E_SHAPE, "couple_inv: bad shape");
  const long long total = (long long)batch * d2;
  couple_inv_kernel<<<bh_blocks(total), kBhThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      state_in, s_raw, t, d, d2, col0, total, state_out, static_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("couple_inv");
}

extern "C" int vub_lstm_input(const float* x_t, int64_t x_pitch, int32_t batch, int32_t n, int32_t xpad, void* cat_bf16,
                              int32_t cat_pitch, int32_t hdim, void* stream) {
  clear_error();
  if (!x_t || !cat_bf16) return fail(VUB_E_NULL, "lstm_input: NULL pointer");
  if (batch <= 0 || n <= 0 || n > xpad || x_pitch < n || cat_pitch < hdim + xpad) return fail(VUB_E_SHAPE, "lstm_input: bad shape");
  const long long total = (long long)batch * xpad;
  lstm_input_kernel<<<bh_blocks(total), kBhThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x_t, x_pitch, n, xpad, total, static_cast<__nv_bfloat16*>(cat_bf16), cat_pitch, hdim);
  return check_launch("lstm_input");
}

extern "C" int vub_lstm_cell(const float* gates, float* c, int32_t batch, int32_t hdim, void* h_lin_bf16, void* h_cat_bf16,
                             int32_t cat_pitch, float* h_f32, void* stream) {
  clear_error();
  if (!gates || !c || !h_cat_bf16) return fail(VUB_E_NULL, "lstm_cell: NULL pointer");
  if (batch <= 0 || hdim <= 0 || cat_pitch < hdim) return fail(VUB_E_SHAPE, "lstm_cell: bad shape");
  const long long total = (long long)batch * hdim;
  lstm_cell_kernel<<<bh_blocks(total), kBhThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      gates, c, hdim, total, static_cast<__nv_bfloat16*>(h_lin_bf16), static_cast<__nv_bfloat16*>(h_cat_bf16), cat_pitch,
      h_f32);
  return check_launch("lstm_cell");
}

extern "C" int vub_decoder_step(float* x, const float* v, int32_t batch, int32_t n, int32_t vpitch, int32_t xpad, int32_t t,
                                int32_t steps, float* xs, float* cs, void* x_cat_bf16, int32_t cat_pitch, int32_t hdim,
                                void* stream) {
  clear_error();
  if (!x || !v || !xs || !cs || !x_cat_bf16) return fail(VUB_E_NULL, "decoder_step: NULL pointer");
  if (batch <= 0 || n <= 0 || n > xpad || vpitch < n || t < 0 || t >= steps || cat_pitch < hdim + xpad)
    return fail(VUB_E_SHAPE, "decoder_step: bad shape");
  const long long total = (long long)batch * xpad;
  decoder_step_kernel<<<bh_blocks(total), kBhThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, v, n, vpitch, xpad, t, steps, total, xs, cs, static_cast<__nv_bfloat16*>(x_cat_bf16), cat_pitch, hdim);
  return check_launch("decoder_step");
}

extern "C" int vub_decoder_init(const float* b_lat, const float* x0, int32_t batch, int32_t hdim, int32_t n, int32_t xpad,
                                float* c, float* x, void* cat_bf16, int32_t cat_pitch, void* stream) {
  clear_error();
  if (!b_lat || !x0 || !c || !x || !cat_bf16) return fail(VUB_E_NULL, "decoder_init: NULL pointer");
  if (batch <= 0 || hdim <= 0 || n <= 0 || n > xpad || cat_pitch != hdim + xpad) return fail(VUB_E_SHAPE, "decoder_init: bad shape");
  const long long total = (long long)batch * cat_pitch;
  decoder_init_kernel<<<bh_blocks(total), kBhThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      b_lat, x0, hdim, n, xpad, batch, c, x, static_cast<__nv_bfloat16*>(cat_bf16), cat_pitch);
  return check_launch("decoder_init");
}
