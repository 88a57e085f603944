// Shared epilogue of the conv kernels: the fused tail of NormConv2d / VunetRNB / Upsample
// (lib/modules.py:144, 225-233, 24-34): y = scale*acc + shift (+ residual) is computed by the caller in
// fp32; this header rounds ONCE per stored tensor and scatters to the requested layout(s).
#pragma once
#include <cuda_bf16.h>

#include "../../include/vunet_b200.h"
#include "ptx.cuh"

namespace vub {

struct EpiParams {
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out_raw;
  __nv_bfloat16* out_elu;
  float* out_f32;
  int hout, wout;
  int cout, cout_valid;
  uint32_t flags;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& a, float& b) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  float2 f = __bfloat1622float2(t);
  a = f.x;
  b = f.y;
}
__device__ __forceinline__ void add_bf16x8(const uint4& u, float* y) {
  float a, b;
  unpack_bf16x2(u.x, a, b); y[0] += a; y[1] += b;
  unpack_bf16x2(u.y, a, b); y[2] += a; y[3] += b;
  unpack_bf16x2(u.z, a, b); y[4] += a; y[5] += b;
  unpack_bf16x2(u.w, a, b); y[6] += a; y[7] += b;
}

// Stores CH consecutive channels [c_global, c_global+CH) of output pixel (n, oh, ow).
template <int CH>
__device__ __forceinline__ void epi_store(const EpiParams& p, float (&y)[CH], int c_global, int n, int oh,
                                          int ow) {
  const long long pix = ((long long)n * p.hout + oh) * p.wout + ow;
  if (p.flags & VUB_CONV_U8_HWC) {
    // video frame: (clamp((y + 1) / 2, 0, 1) * 255) truncated to uint8, HWC — same fp32 arithmetic as
    // frames_to_uint8_kernel on the fp32 image, so the bytes are identical
    uint8_t* o = reinterpret_cast<uint8_t*>(p.out_f32) + pix * p.cout_valid;
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (c_global + i < p.cout_valid) {
        const float v = fminf(fmaxf((y[i] + 1.0f) / 2.0f, 0.0f), 1.0f) * 255.0f;
        o[c_global + i] = static_cast<uint8_t>(static_cast<int>(v));
      }
    return;
  }
  if (p.flags & VUB_CONV_F32_NCHW) {
    const long long plane = (long long)p.hout * p.wout;
    float* o = p.out_f32 + ((long long)n * p.cout_valid) * plane + (long long)oh * p.wout + ow;
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (c_global + i < p.cout_valid) o[(long long)(c_global + i) * plane] = y[i];
    return;
  }
  if (p.flags & VUB_CONV_F32_NHWC) {
    float4* o = reinterpret_cast<float4*>(p.out_f32 + pix * p.cout + c_global);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
    return;
  }
  long long obase;
  if (p.flags & VUB_CONV_D2S) {
    // DepthToSpace(2), DCR order: out[n, 2h+b1, 2w+b2, c'] = conv[n, h, w, (2*b1+b2)*C' + c']
    const int cq = p.cout >> 2;
    const int qd = c_global / cq;
    const int cp = c_global - qd * cq;
    const long long opix =
        ((long long)n * (2 * p.hout) + 2 * oh + (qd >> 1)) * (2 * p.wout) + 2 * ow + (qd & 1);
    obase = opix * cq + cp;
  } else {
    obase = pix * p.cout + c_global;
  }
  if (p.out_raw != nullptr) {
    uint4* o = reinterpret_cast<uint4*>(p.out_raw + obase);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j)
      o[j] = make_uint4(pack_bf16x2(y[8 * j + 0], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                        pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
  }
  if (p.out_elu != nullptr) {
    const int mode = act_mode_of(p.flags);
#pragma unroll
    for (int i = 0; i < CH; ++i) y[i] = act_apply(y[i], mode);
    uint4* o = reinterpret_cast<uint4*>(p.out_elu + obase);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j)
      o[j] = make_uint4(pack_bf16x2(y[8 * j + 0], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                        pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
  }
}

// ------------------------------------------------------------------------------------------------
// Coalesced bf16 stores through a per-warp shared-memory staging tile.
//
// After tcgen05.ld a lane owns ONE pixel and all its channels, so a direct `st.global.v4` per lane hits
// 32 different cache lines per warp instruction and the L1 tag stage (1 line/cycle) becomes the
// bottleneck of every HBM-bound layer (ncu: profiles/r01_ncu_notes.md).  Instead each warp parks its
// 32 pixels x GW channels in shared memory and writes them back with consecutive lanes covering
// consecutive 16-byte chunks, i.e. 512 contiguous bytes per warp instruction.
//
// Layout: unpadded rows of GW*2 bytes with an XOR swizzle of the 16-byte chunk index, chosen so that BOTH
// access patterns are bank-conflict free for 128-bit accesses (a quarter warp = 8 lanes must touch 8
// different 16-byte bank groups): "row per lane" (lane r, fixed chunk) and "linear" (8 consecutive chunks).
// The earlier +16-byte row pad made only the first pattern conflict free; the linear pass took 2 wavefronts
// per quarter warp, and ncu showed the shared-memory data pipe (LSU wavefronts + the tensor core's operand
// reads) saturated in the row-streaming kernel (profiles/r01_ncu_rows_smem.txt).
// ------------------------------------------------------------------------------------------------
template <int GW>
struct EpiStage {
  static constexpr int CPR = GW / 8;                 // 16-byte chunks per staged row
  static constexpr int PITCH = GW * 2;
  static constexpr int TILE_BYTES = 32 * PITCH;      // one staged copy (raw OR elu) of a warp's tile
  static constexpr int WARP_BYTES = 2 * TILE_BYTES;  // raw + elu
  // byte offset of 16-byte chunk `chunk` of staged row `row`
  __device__ __forceinline__ static uint32_t off(int row, int chunk) {
    if constexpr (CPR >= 8) {
      return row * PITCH + ((chunk ^ (row & 7)) << 4);
    } else {
      constexpr int RP = 8 / CPR;  // rows per 128 bytes
      return row * PITCH + ((chunk ^ ((row / RP) & (CPR - 1))) << 4);
    }
  }
};

// Staging layout for tiles that a TMA tensor-map copy writes or reads (cp.async.bulk.tensor with 128-byte swizzle):
// the box row is 128 bytes — one pixel at GW = 64, a PAIR of pixels at GW = 32 (a tensor box moves 3 cycles per box
// row whatever its width, profiles/r01_exp_rates_tma.log, so 64-byte rows would halve the copy rate) — and the
// hardware XORs the 16-byte chunk index with (box row & 7).  Tiles must start on a 1024-byte boundary.  Both access
// patterns of the epilogue stay bank-conflict free: "pixel per lane, fixed chunk" hits 8 different 16-byte columns
// per quarter warp (rows r, r+1, r+2, r+3 x the two pixel halves for GW = 32).
template <int GW>
struct TmaStage {
  static constexpr int CPR = GW / 8;
  static constexpr int TILE_BYTES = 32 * GW * 2;
  static constexpr int BOX_ROWS = GW == 32 ? 16 : 32;   // 128-byte rows per 32-pixel tile
  __device__ __forceinline__ static uint32_t off(int pix, int chunk) {
    static_assert(GW == 32 || GW == 64, "TMA staging covers 32- and 64-channel tiles");
    if constexpr (GW == 64) {
      return pix * 128 + ((chunk ^ (pix & 7)) << 4);
    } else {
      const int row = pix >> 1;
      return row * 128 + (((((pix & 1) << 2) | chunk) ^ (row & 7)) << 4);
    }
  }
};

// Stage 32 consecutive channels of this lane's pixel into a TMA-layout tile (see TmaStage).
template <int GW>
__device__ __forceinline__ void stage_put32_tma(uint32_t tile, int lane, int c, const float (&y)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(tile + TmaStage<GW>::off(lane, (c >> 3) + j),
           make_uint4(pack_bf16x2(y[8 * j + 0], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                      pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7])));
}

// Stage 32 consecutive channels (starting at channel `c` of the group) of this lane's pixel.
// `tile` is a shared-space byte address.
template <int GW>
__device__ __forceinline__ void stage_put32(uint32_t tile, int lane, int c, const float (&y)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(tile + EpiStage<GW>::off(lane, (c >> 3) + j),
           make_uint4(pack_bf16x2(y[8 * j + 0], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                      pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7])));
}

// Element offset (in `out`) of channel `c_global` of output pixel (n, oh, ow); honours VUB_CONV_D2S.
__device__ __forceinline__ long long epi_out_offset(const EpiParams& p, int c_global, int n, int oh, int ow) {
  if (p.flags & VUB_CONV_D2S) {
    const int cq = p.cout >> 2;
    const int qd = c_global / cq;
    const int cp = c_global - qd * cq;
    const long long opix =
        ((long long)n * (2 * p.hout) + 2 * oh + (qd >> 1)) * (2 * p.wout) + 2 * ow + (qd & 1);
    return opix * cq + cp;
  }
  return (((long long)n * p.hout + oh) * p.wout + ow) * p.cout + c_global;
}

// Write a staged tile back: `coord(row, n, oh, ow)` yields the pixel of staged row `row` and returns
// false for rows that must not be stored (batch tail).  `c_group` = first channel of the staged group.
template <int GW, class CoordFn>
__device__ __forceinline__ void stage_flush(uint32_t tile, __nv_bfloat16* out, const EpiParams& p, int lane,
                                            int c_group, CoordFn coord) {
  constexpr int CPR = GW / 8;  // 16-byte chunks per staged row
#pragma unroll
  for (int k = 0; k < CPR; ++k) {
    const int c = lane + 32 * k;
    const int row = c / CPR;
    const int col = c - row * CPR;
    int n, oh, ow;
    if (coord(row, n, oh, ow)) {
      const uint4 v = lds128(tile + EpiStage<GW>::off(row, col));
      *reinterpret_cast<uint4*>(out + epi_out_offset(p, c_group + col * 8, n, oh, ow)) = v;
    }
  }
}

// Same, when the warp's 32 pixels x GW channels form ONE contiguous block of global memory (the
// row-streaming kernel without DepthToSpace): a straight linear copy, no per-chunk address math.
template <int GW>
__device__ __forceinline__ void stage_flush_linear(uint32_t tile, __nv_bfloat16* out_block, int lane) {
  constexpr int CPR = GW / 8;
  uint4* dst = reinterpret_cast<uint4*>(out_block);
#pragma unroll
  for (int k = 0; k < CPR; ++k) {
    const int c = lane + 32 * k;
    const int row = c / CPR;
    const int col = c - row * CPR;
    dst[c] = lds128(tile + EpiStage<GW>::off(row, col));
  }
}

// Inverse direction for the residual: coalesced global -> staging (linear chunks), then each lane
// reads its own pixel row.
template <int GW>
__device__ __forceinline__ void stage_fill_linear(uint32_t tile, const uint4 (&chunks)[GW / 8], int lane) {
  constexpr int CPR = GW / 8;
#pragma unroll
  for (int k = 0; k < CPR; ++k) {
    const int c = lane + 32 * k;
    const int row = c / CPR;
    const int col = c - row * CPR;
    sts128(tile + EpiStage<GW>::off(row, col), chunks[k]);
  }
}

}  // namespace vub
