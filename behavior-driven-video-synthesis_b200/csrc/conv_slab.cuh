// Host-side description of one launch of the activation-slab kernel (conv_slab.cu): a 3x3 stride-1 conv
// (flip = 0) or its data gradient (flip = 1) over 64-channel blocks of one or two NHWC sources.
#pragma once
#include <cuda_runtime.h>

#include "conv_tiles.cuh"

namespace vub {

struct SlabJob {
  const void* x[2];      // A sources, NHWC bf16 [batch, h, w, c_src]
  int nsrc, c_src;       // c_src multiple of 64
  int batch, h, w;       // w in {32, 64}; h multiple of 256 / w
  const void* w_packed;  // B: [n_total rows][ktot] bf16, K order (source, tap, channel)
  const void* w_tiled;   // optional tile-major pre-swizzled copy [ktot/64][w_rows][64], else nullptr
  int w_rows;
  long long ktot;
  int n_total;           // multiple of 128
  int flip;              // 0 forward, 1 data gradient (taps mirrored)
  int n_split;           // as in TileJob
  TileEpi epi[2];
};

bool conv_slab_shape_ok(int c_src, int h, int w, int n_total, int k_total, long long pix_tiles);
int conv_slab_launch(const SlabJob& job, cudaStream_t stream);

}  // namespace vub
