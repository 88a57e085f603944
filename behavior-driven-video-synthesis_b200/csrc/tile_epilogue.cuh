// Mode-driven epilogue shared by the persistent tile kernel (conv_tiles.cu) and the slab kernel
// (conv_slab.cu): one warp drains its 32 TMEM lanes (= 32 pixels) x BN accumulator columns.
//   y = scale * acc + shift;  y *= ELU'(mask) (data gradients);  y += addend (+ addend2);
//   bf16 raw and/or ELU(y) outputs staged through shared memory (coalesced 16-byte chunks), DepthToSpace or
//   strided pixel scatter, or fp32 NHWC / NCHW outputs.
// Fuses lib/modules.py:144, 225-233, 24-34 (forward) and their autograd counterparts (backward).
#pragma once
#include "conv_epilogue.cuh"
#include "conv_tiles.cuh"
#include "ptx.cuh"

namespace vub {

struct EpiGeom {
  int tw, th, tw_log2, th_log2;  // a 128-pixel tile = tw x th pixels of 128 / (tw * th) images
  int batch;
};

__device__ __forceinline__ long long tile_out_offset(const TileEpi& e, int col, int n, int oh, int ow) {
  if (e.flags & VUB_CONV_F32_NCHW)
    return (((long long)n * e.n_valid + col) * e.out_h + oh) * e.out_w + ow;
  if (e.flags & VUB_CONV_D2S) {
    const int qd = col / e.out_c;
    const int cp = col - qd * e.out_c;
    const long long opix = ((long long)n * e.out_h + 2 * oh + (qd >> 1)) * e.out_w + 2 * ow + (qd & 1);
    return opix * e.out_c + cp;
  }
  const long long opix =
      ((long long)n * e.out_h + oh * e.pix_stride + e.pix_off_h) * e.out_w + ow * e.pix_stride + e.pix_off_w;
  return opix * e.out_c + e.c_off + col;
}

// `trow`: TMEM address of this warp's lane quarter at the accumulator's first column; `t0/t1/t2`: this warp's
// staging tiles (EpiStage<GW>::TILE_BYTES each; t2 only when an addend2 is present); `s_scale/s_shift`: already
// offset to the tile's first GEMM column (both NULL = identity); `ecol`: the tile's first column as the epilogue `e` numbers it;
// (n0, h0, w0): tile origin.  Arrives once (lane 0) on `acc_empty` (unless NULL) as soon as the accumulator has
// been read.  With `acc_full` given, the function itself waits for the accumulator — AFTER it has put the first
// channel group's input loads (addend / mask / addend2) in flight, which do not depend on it.
template <int BN, int GW>
__device__ __forceinline__ void tile_epilogue_staged(const TileEpi& e, const EpiGeom gm, const float* s_scale,
                                                     const float* s_shift, uint32_t trow, uint32_t t0, uint32_t t1,
                                                     uint32_t t2, int lane, int q, int ecol, int n0, int h0, int w0,
                                                     uint64_t* acc_empty, uint64_t* acc_full = nullptr,
                                                     uint32_t acc_parity = 0) {
  constexpr int CPR = GW / 8;
  const int m = q * 32 + lane;
  const int n = n0 + (m >> (gm.tw_log2 + gm.th_log2));
  const int oh = h0 + ((m >> gm.tw_log2) & (gm.th - 1));
  const int ow = w0 + (m & (gm.tw - 1));
  const bool valid = n < gm.batch;
  const bool f32_out = (e.flags & (VUB_CONV_F32_NHWC | VUB_CONV_F32_NCHW)) != 0;
  // pixel of staged row `row` of this warp
  auto row_pixel = [&](int row, int& rn, int& roh, int& row_w) {
    const int mm = q * 32 + row;
    rn = n0 + (mm >> (gm.tw_log2 + gm.th_log2));
    roh = h0 + ((mm >> gm.tw_log2) & (gm.th - 1));
    row_w = w0 + (mm & (gm.tw - 1));
    return rn < gm.batch;
  };
  // All loads of a tensor are issued before the first shared-memory store: sts128 is a volatile asm with a
  // memory clobber, so a load -> store -> load -> store sequence serialises CPR global-memory latencies
  // (measured: the ELU' mask alone doubled the data-gradient time of the HBM-bound layers).
  auto stage_fill = [&](uint32_t tile_addr, const __nv_bfloat16* src, int col0) {
    uint4 v[CPR];
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
      const int c = lane + 32 * k;
      const int row = c / CPR;
      const int col = c - row * CPR;
      int rn, roh, row_w;
      v[k] = make_uint4(0u, 0u, 0u, 0u);
      if (row_pixel(row, rn, roh, row_w))
        v[k] = *reinterpret_cast<const uint4*>(src + tile_out_offset(e, col0 + col * 8, rn, roh, row_w));  // may alias the output
    }
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
      const int c = lane + 32 * k;
      const int row = c / CPR;
      const int col = c - row * CPR;
      sts128(tile_addr + EpiStage<GW>::off(row, col), v[k]);
    }
  };
  auto stage_flush_out = [&](uint32_t tile_addr, __nv_bfloat16* dst, int col0) {
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
      const int c = lane + 32 * k;
      const int row = c / CPR;
      const int col = c - row * CPR;
      int rn, roh, row_w;
      if (row_pixel(row, rn, roh, row_w))
        *reinterpret_cast<uint4*>(dst + tile_out_offset(e, col0 + col * 8, rn, roh, row_w)) =
            lds128(tile_addr + EpiStage<GW>::off(row, col));
    }
  };

#pragma unroll 1
  for (int g0 = 0; g0 < BN; g0 += GW) {
    const int col0 = ecol + g0;  // first (epilogue-local) column of this group
    uint4 va[4], vm[4], vb[4];   // GW == 32: this lane's addend / mask / addend2 values
    if (!f32_out) {
      if constexpr (GW == 32) {
        // HBM-bound layers: every lane loads ITS OWN pixel's 32 channels of all input tensors straight into
        // registers, all loads in flight before the first use.  No shared-memory staging for inputs: these kernels
        // are bound by the shared-memory data pipe (tensor-core operand reads + staging), and the warp's four
        // 16-byte loads per tensor still cover whole sectors, so L2 traffic is unchanged.
#pragma unroll
        for (int j = 0; j < 4; ++j) va[j] = vm[j] = vb[j] = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long off = tile_out_offset(e, col0 + j * 8, n, oh, ow);
            if (e.addend != nullptr) va[j] = *reinterpret_cast<const uint4*>(e.addend + off);  // may alias the output
            if (e.mask_elu != nullptr) vm[j] = *reinterpret_cast<const uint4*>(e.mask_elu + off);
            if (e.addend2 != nullptr) vb[j] = *reinterpret_cast<const uint4*>(e.addend2 + off);
          }
        }
      } else {
        if (e.addend != nullptr) stage_fill(t0, e.addend, col0);
        if (e.mask_elu != nullptr) stage_fill(t1, e.mask_elu, col0);
        if (e.addend2 != nullptr) stage_fill(t2, e.addend2, col0);
      }
      __syncwarp();
    }
    if (g0 == 0 && acc_full != nullptr) {
      mbar_wait(acc_full, acc_parity);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = g0; c0 < g0 + GW; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(trow + c0, acc);
      tmem_ld_wait();
      if (c0 + 32 == BN) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && acc_empty != nullptr) mbar_arrive(acc_empty);
      }
      float y[32];
      if (s_scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = fmaf(__uint_as_float(acc[i]), s_scale[c0 + i], s_shift[c0 + i]);
      } else {  // identity affine (data gradients): no broadcast loads from shared memory
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(acc[i]);
      }
      if (f32_out) {
        if (valid) {
          if (e.flags & VUB_CONV_F32_NCHW) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (ecol + c0 + i < e.n_valid)
                e.out_f32[tile_out_offset(e, ecol + c0 + i, n, oh, ow)] = y[i];
          } else {
            float4* o = reinterpret_cast<float4*>(e.out_f32 + tile_out_offset(e, ecol + c0, n, oh, ow));
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
          }
        }
        continue;
      }
      const int ch0 = (c0 - g0) >> 3;  // first 16-byte chunk of these 32 channels in this lane's staged row
      if (e.mask_elu != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float ev[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if constexpr (GW == 32) add_bf16x8(vm[j], ev);
          else add_bf16x8(lds128(t1 + EpiStage<GW>::off(lane, ch0 + j)), ev);
#pragma unroll
          for (int i = 0; i < 8; ++i) y[8 * j + i] *= act1_grad(ev[i], (e.flags & VUB_CONV_RELU) != 0);
        }
      }
      if (e.addend != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (GW == 32) add_bf16x8(va[j], &y[8 * j]);
          else add_bf16x8(lds128(t0 + EpiStage<GW>::off(lane, ch0 + j)), &y[8 * j]);
        }
      }
      if (e.addend2 != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (GW == 32) add_bf16x8(vb[j], &y[8 * j]);
          else add_bf16x8(lds128(t2 + EpiStage<GW>::off(lane, ch0 + j)), &y[8 * j]);
        }
      }
      if (e.out_raw != nullptr) stage_put32<GW>(t0, lane, c0 - g0, y);
      if (e.out_elu != nullptr) {
        const bool relu = (e.flags & VUB_CONV_RELU) != 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = act1(y[i], relu);
        stage_put32<GW>(t1, lane, c0 - g0, y);
      }
    }
    if (!f32_out) {
      __syncwarp();
      if (e.out_raw != nullptr) stage_flush_out(t0, e.out_raw, col0);
      if (e.out_elu != nullptr) stage_flush_out(t1, e.out_elu, col0);
      __syncwarp();
    }
  }
}

// The same epilogue for the common case that the tile's 128 pixels are ONE contiguous run of the NHWC output
// (full-width tiles, stride 1, no DepthToSpace, bf16 outputs): every address is `base + row * out_c + column`,
// no per-chunk (n, h, w) decoding and no 64-bit multiplies.  The generic version above executed ~4500 instructions
// per warp and tile and the kernels built on it (10 000 SASS instructions) spent 24 % of their warp stalls on
// instruction fetch (ncu, 128->128 @64^2); it stays for the scattering / fp32 layouts.
// `pix0`: index of this warp's first pixel in the output tensor; `n_valid`: pixels of this warp inside the tensor.
// With VUB_CONV_D2S (`d2s_w` = width of the conv's own output grid, the warp's 32 pixels inside one image row, GW
// channels inside one quadrant): the quadrant's pixels are every second pixel of one row of the shuffled tensor, so
// the addressing stays linear with twice the row pitch.
template <int BN, int GW>
__device__ __forceinline__ void tile_epilogue_contig(const TileEpi& e, const float* s_scale, const float* s_shift,
                                                     uint32_t trow, uint32_t t0, uint32_t t1, uint32_t t2, int lane,
                                                     int ecol, long long pix0, int n_valid, uint64_t* acc_empty,
                                                     uint64_t* acc_full, uint32_t acc_parity, int d2s_w = 0,
                                                     uint32_t acc_empty_cluster = 0) {
  // acc_empty_cluster: shared::cluster address of the accumulator-empty barrier of a CTA pair's leader (conv_pair)
  constexpr int CPR = GW / 8;
  const bool d2s = d2s_w > 0;
  const bool valid = lane < n_valid;
  long long d2s_row = 0;  // D2S: index of output pixel (2h, 2w) of this warp's first pixel
  if (d2s) {
    const long long rowi = pix0 / d2s_w;  // n * H + h
    const int w = static_cast<int>(pix0 - rowi * d2s_w);
    d2s_row = (rowi * 2) * (2LL * d2s_w) + 2 * w;
  }
#pragma unroll 1
  for (int g0 = 0; g0 < BN; g0 += GW) {
    // element offset of (first pixel, first column of this channel group) and the pitch between the warp's pixels
    long long base;
    int oc;
    if (!d2s) {
      oc = e.out_c;
      base = pix0 * e.out_c + e.c_off + ecol;  // (g0 is added where the group is addressed)
    } else {
      const int col = ecol + g0;
      const int qd = col / e.out_c, cp = col - qd * e.out_c;
      oc = 2 * e.out_c;
      base = (d2s_row + (qd >> 1) * (2LL * d2s_w) + (qd & 1)) * e.out_c + cp - g0;
    }
    uint4 va[4], vm[4], vb[4];  // GW == 32: this lane's addend / mask / addend2 values
    if constexpr (GW == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) va[j] = vm[j] = vb[j] = make_uint4(0u, 0u, 0u, 0u);
      if (valid) {
        const long long off = base + lane * oc + g0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (e.addend != nullptr) va[j] = *reinterpret_cast<const uint4*>(e.addend + off + 8 * j);  // may alias the output
          if (e.mask_elu != nullptr) vm[j] = *reinterpret_cast<const uint4*>(e.mask_elu + off + 8 * j);
          if (e.addend2 != nullptr) vb[j] = *reinterpret_cast<const uint4*>(e.addend2 + off + 8 * j);
        }
      }
    } else {
      // staged inputs: all loads of a tensor in flight before its first shared-memory store
      auto fill = [&](uint32_t tile_addr, const __nv_bfloat16* src) {
        uint4 v[CPR];
#pragma unroll
        for (int k = 0; k < CPR; ++k) {
          const int c = lane + 32 * k;
          const int row = c / CPR, col = c - row * CPR;
          v[k] = make_uint4(0u, 0u, 0u, 0u);
          if (row < n_valid) v[k] = *reinterpret_cast<const uint4*>(src + base + row * oc + g0 + col * 8);
        }
#pragma unroll
        for (int k = 0; k < CPR; ++k) {
          const int c = lane + 32 * k;
          const int row = c / CPR, col = c - row * CPR;
          sts128(tile_addr + EpiStage<GW>::off(row, col), v[k]);
        }
      };
      if (e.addend != nullptr) fill(t0, e.addend);
      if (e.mask_elu != nullptr) fill(t1, e.mask_elu);
      if (e.addend2 != nullptr) fill(t2, e.addend2);
      __syncwarp();
    }
    if (g0 == 0 && acc_full != nullptr) {
      mbar_wait(acc_full, acc_parity);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = g0; c0 < g0 + GW; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(trow + c0, acc);
      tmem_ld_wait();
      if (c0 + 32 == BN) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && acc_empty != nullptr) mbar_arrive(acc_empty);
        if (lane == 0 && acc_empty_cluster != 0) mbar_arrive_cluster(acc_empty_cluster);
      }
      float y[32];
      if (s_scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = fmaf(__uint_as_float(acc[i]), s_scale[c0 + i], s_shift[c0 + i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(acc[i]);
      }
      const int ch0 = (c0 - g0) >> 3;
      if (e.mask_elu != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float ev[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if constexpr (GW == 32) add_bf16x8(vm[j], ev);
          else add_bf16x8(lds128(t1 + EpiStage<GW>::off(lane, ch0 + j)), ev);
#pragma unroll
          for (int i = 0; i < 8; ++i) y[8 * j + i] *= act1_grad(ev[i], (e.flags & VUB_CONV_RELU) != 0);
        }
      }
      if (e.addend != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (GW == 32) add_bf16x8(va[j], &y[8 * j]);
          else add_bf16x8(lds128(t0 + EpiStage<GW>::off(lane, ch0 + j)), &y[8 * j]);
        }
      }
      if (e.addend2 != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (GW == 32) add_bf16x8(vb[j], &y[8 * j]);
          else add_bf16x8(lds128(t2 + EpiStage<GW>::off(lane, ch0 + j)), &y[8 * j]);
        }
      }
      if (e.out_raw != nullptr) stage_put32<GW>(t0, lane, c0 - g0, y);
      if (e.out_elu != nullptr) {
        const bool relu = (e.flags & VUB_CONV_RELU) != 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = act1(y[i], relu);
        stage_put32<GW>(t1, lane, c0 - g0, y);
      }
    }
    __syncwarp();
    auto flush = [&](uint32_t tile_addr, __nv_bfloat16* dst) {
#pragma unroll
      for (int k = 0; k < CPR; ++k) {
        const int c = lane + 32 * k;
        const int row = c / CPR, col = c - row * CPR;
        if (row < n_valid)
          *reinterpret_cast<uint4*>(dst + base + row * oc + g0 + col * 8) = lds128(tile_addr + EpiStage<GW>::off(row, col));
      }
    };
    if (e.out_raw != nullptr) flush(t0, e.out_raw);
    if (e.out_elu != nullptr) flush(t1, e.out_elu);
    __syncwarp();
  }
}

}  // namespace vub
