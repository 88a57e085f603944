// Host-side description of one launch of the persistent tile kernel (conv_tiles.cu).  Forward convs,
// data gradients of stride-1 / stride-2 / DepthToSpace convs are all expressed as a list of k-steps
// (which tensor map, which pixel shift, which channel block, which weight column block) plus an
// epilogue description, so one kernel serves them all.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace vub {

constexpr int kTileMaxKSteps = 80;

struct KStep {
  uint8_t map;     // index into the A tensor maps (source tensor or parity sub-grid)
  int8_t dh, dw;   // pixel shift of the A box relative to the output tile origin
  uint8_t cblk;    // channel block (units of BKC) inside the A tensor
  uint16_t bblk;   // column block (units of BKC) inside the packed weight matrix
  uint16_t pad;
};

// one A source: NHWC bf16 tensor (or a strided parity view of one)
struct TileSrc {
  const void* base;
  int channels;        // innermost extent
  int w, h;            // extents of the (sub-)grid the tile coordinates address
  long long stride_w;  // element strides
  long long stride_h;
  long long stride_n;
};

struct TileEpi {
  const float* scale;            // per GEMM column, or nullptr (identity)
  const float* shift;            // per GEMM column, or nullptr (zero)
  const __nv_bfloat16* addend;   // added after scale/shift; same addressing as the bf16 outputs; or nullptr
  const __nv_bfloat16* addend2;  // second addend (dgrad: residual identity path + earlier skip contributions)
  const __nv_bfloat16* mask_elu; // dgrad: y *= ELU'(x) reconstructed from the stored ELU(x): e > 0 ? 1 : e + 1
  __nv_bfloat16* out_raw;
  __nv_bfloat16* out_elu;
  float* out_f32;
  int out_h, out_w;              // extents of the output tensor addressed by the epilogue
  int out_c;                     // channels per output pixel (row pitch of the bf16 outputs)
  int c_off;                     // channel offset of GEMM column 0 inside an output pixel
  int pix_stride, pix_off_h, pix_off_w;  // output pixel = (oh*pix_stride + pix_off_h, ow*pix_stride + pix_off_w)
  int n_valid;                   // GEMM columns actually stored (F32_NCHW only)
  uint32_t flags;                // VUB_CONV_D2S | VUB_CONV_F32_NHWC | VUB_CONV_F32_NCHW
};

struct TileJob {
  TileSrc src[4];
  int n_src;
  int bkc;               // 32 or 64
  const void* w;         // packed weights [n_total rows][ktot] bf16, K-major
  const void* w_tiled;   // optional tile-major pre-swizzled copy [ktot/64][w_rows][64] (bkc == 64), else nullptr
  int w_rows;
  long long ktot;        // row pitch of w in elements
  int n_total;           // GEMM N (rows of w used); multiple of the N tile
  int bn;                // N tile override (0 = min(n_total, 256))
  int n_split;           // GEMM columns [0, n_split) use epi[0], the rest epi[1] with columns renumbered from 0
                         // (0 = no split; must be a multiple of the N tile).  scale/shift always come from epi[0]
                         // and are indexed by the global GEMM column.
  int batch, hout, wout; // tile grid: output pixels enumerated (n, oh, ow) over hout x wout
  KStep ksteps[kTileMaxKSteps];
  int n_ksteps;
  TileEpi epi[2];
};

int conv_tiles_launch(const TileJob& job, cudaStream_t stream);

}  // namespace vub
