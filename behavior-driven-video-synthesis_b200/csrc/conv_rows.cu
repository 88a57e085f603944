// Row-streaming forward convolution for the wide, thin layers of the VUnet (output width >= 128,
// <= 64 channels per source, <= 128 output channels; stride 1; 3x3 or 1x1) — the layers that are
// HBM-bound or nearly so (32->32 @256^2, 64->64 @128^2, the skip `nin`s, ups.5,
// out_conv).
//
// One CTA owns a strip of TH output rows x 128 output columns of one image:
//   * the packed weights of ALL taps stay resident in shared memory for the CTA's lifetime;
//   * every input row (130 pixels incl. the conv halo, out-of-bounds zero-filled by TMA) is loaded from
//     L2/HBM ONCE into a ring of row buffers and feeds the 3 output rows that need it;
//   * the 9 filter taps are 9 UMMA descriptors into those buffers — a tap's column shift is a start
//     address offset of s * row_bytes into the swizzled buffer (verified on hardware:
//     profiles/r01_exp_rowshift.log), so no im2col copy exists in shared memory either;
//   * accumulators are double-buffered in TMEM; two epilogue warpgroups alternate output rows, each
//     prefetching its next row's residual while the tensor core works on the other buffer.
// L2->SM traffic per output row drops from 9 (+ weights) row-reads in the per-tap kernel to
// (TH+2)/TH row-reads, which is what lets these layers approach the HBM roofline.
//
// Replaces the same reference code as conv_fwd.cu (lib/modules.py:140-145, 221-233, 24-34).
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.h"
#include "conv_epilogue.cuh"
#include "conv_tiles.cuh"
#include "ptx.cuh"
#include "tile_epilogue.cuh"

namespace vub {

constexpr int kRowsThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue A, warps 6-9 epilogue B
constexpr int kRowsThreadsElu = 448;  // + warps 10-13: ELU-on-load transform of freshly landed input rows
constexpr int kRowsMaxRing = 8;
constexpr int kRowsW = 128;  // output columns per CTA = UMMA M

struct RowsKParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  EpiParams ep;
  int batch;
  int nsrc, ntaps;
  int strips_w;         // wout / 128
  long long total_rows; // strips_w * batch * hout: output rows of 128 pixels, the unit of work
  int ring;             // row-buffer ring slots
  // data-gradient mode (vub_conv_dgrad on the wide, thin layers): taps mirrored, mode-driven shared epilogue
  // (ELU' mask, two addends); with two halves the accumulator columns [0, BN/2) and [BN/2, BN) go to the two
  // sources of a concat.
  int flip;
  int n_halves;         // 0 = forward epilogue (ep), 1 or 2 = data-gradient epilogue (bwd[])
  int stage_tiles;      // staging tiles per epilogue warp in data-gradient mode
  TileEpi bwd[2];
  // ELU-on-load (VUB_CONV_ELU_IN0/1): bit s set = source s holds RAW activations; four extra warps apply ELU to
  // every input row in place in its ring slot before the tensor core reads it, so the producer of that tensor
  // never writes an ELU copy to HBM and the residual (the same raw tensor) is an L2 hit.
  int elu_in;
  // A operand in tensor memory (3x3 only): every input row is copied shared -> tensor memory three times (one
  // copy per column tap, tcgen05.cp from a shifted start row) when it arrives and the nine taps' MMAs read A from
  // there, so the tensor core fetches an input row from shared memory 3 times instead of 9 — the shared-memory
  // data pipe is what bounds this kernel (profiles/r01_ncu_rows_smem.txt).
  int atm;
  // Kernel rows as N blocks (3x3, BN 32 / 64): an input row is multiplied ONCE, by the three kernel rows' weight
  // tiles stacked along N ([kh=2 | kh=1 | kh=0] -> output rows r-1, r, r+1 whose accumulators are neighbours in a
  // ring of tensor-memory slots).  6 (12) MMAs of N = 96 (192) per input row instead of 18 (36) of N = 32 (64): an
  // M128 N32 K16 MMA costs 40 cycles for 16 cycles of math because its shared-memory operand fetch is (128 + N) rows
  // per instruction (profiles/r01_exp_rates_mma.log) — stacking N amortises the 128 A rows over three outputs.
  int nblk;
  // ELU-on-load straight into TENSOR memory (nblk kernels): the four transform warps read a landed row from shared
  // memory once, apply ELU in registers and write the activated row — and its two column-shifted copies, built with
  // warp shuffles — into a tensor-memory operand slot (tcgen05.st); the nblk MMAs take A from there.  The activated
  // row never returns to shared memory and is never read from it by the tensor core: per 128-pixel row at 32
  // channels the shared-memory pipe carries ~350 LSU + ~170 tensor-core wavefronts instead of 744 + 414
  // (profiles/r02_ncu_rows.txt) — that pipe, not HBM, bounded the kernel.
  int tst;
  // TMA epilogue (forward, one bf16 output, BN 32 / 64): the finished 32-pixel tiles leave shared memory through
  // tensor-map stores issued by one lane per warp (tma = 1), and the residual rows arrive the same way (tma = 2),
  // instead of LDS + STG / LDG + STS passes of all 32 lanes — the kernel is bound by the shared-memory data pipe
  // (LSU wavefronts, profiles/r01_ncu_rows_smem.txt), which these copies bypass.
  int tma;
  CUtensorMap tmO, tmR;
};

template <int BKC>
struct RowsGeom {
  static constexpr int ROWB = BKC * 2;
  // 130 rows are written per source; buffers are padded to keep every buffer 1024B aligned
  static constexpr int SRC_BYTES = (BKC == 64) ? 136 * 128 : 144 * 64;
};

// Walks the CTA's contiguous range [g, g_end) of output rows.  Row index g = ((strip_w * B + n) * H + h);
// a "segment" is a maximal run of rows inside one image (the unit that shares halo rows).
struct RowWalker {
  long long g, g_end;
  int H, B;
  int n, h, w0;   // current row
  int seg_left;   // rows left in the current segment, including the current one
  __device__ __forceinline__ void init(long long g0, long long g1, int H_, int B_) {
    g = g0;
    g_end = g1;
    H = H_;
    B = B_;
    decode();
  }
  __device__ __forceinline__ void decode() {
    if (g >= g_end) {
      seg_left = 0;
      return;
    }
    h = static_cast<int>(g % H);
    const long long t = g / H;
    n = static_cast<int>(t % B);
    w0 = static_cast<int>(t / B) * 128;
    const long long left = g_end - g;
    seg_left = (H - h) < left ? (H - h) : static_cast<int>(left);
  }
  __device__ __forceinline__ bool valid() const { return g < g_end; }
  __device__ __forceinline__ void next_row() {
    ++g;
    ++h;
    if (--seg_left == 0) decode();
  }
  __device__ __forceinline__ void next_segment() {
    g += seg_left;
    decode();
  }
};

// ELU on 8 packed bf16 values (fp32 math, one MUFU.EX2 each)
__device__ __forceinline__ void elu_bf16x8(uint4& v) {
  float f[8];
  unpack_bf16x2(v.x, f[0], f[1]);
  unpack_bf16x2(v.y, f[2], f[3]);
  unpack_bf16x2(v.z, f[4], f[5]);
  unpack_bf16x2(v.w, f[6], f[7]);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = elu1(f[k]);
  v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// MODE: 0 = forward, 1 = forward with ELU-on-load (4 extra transform warps), 2 = data gradient.  Separate
// instantiations keep each kernel's code small: with both epilogues in one kernel (~9000 SASS instructions) ncu
// showed 16 % of the warp stalls waiting for instruction fetch.
constexpr int kRowsFwd = 0, kRowsFwdElu = 1, kRowsBwd = 2;

template <int BN, int BKC, int NTAPS, int NSRC, int MODE>
__global__ void __launch_bounds__(MODE == kRowsFwdElu ? kRowsThreadsElu : kRowsThreads, 1)
    conv_rows_kernel(const __grid_constant__ RowsKParams p) {
  constexpr bool ELU_IN = MODE == kRowsFwdElu;
  constexpr bool BWD = MODE == kRowsBwd;
  using G = RowsGeom<BKC>;
  constexpr int ROWB = G::ROWB;
  constexpr int SRC_BYTES = G::SRC_BYTES;
  constexpr int WT_BYTES = BN * ROWB;  // one (source, tap) weight tile
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  constexpr int CH = BN >= 32 ? 32 : 16;
  constexpr int NCHUNK = BN / CH;
  constexpr int halo = NTAPS == 9 ? 1 : 0;
  constexpr int kdim = NTAPS == 9 ? 3 : 1;
  constexpr int slot_bytes = NSRC * SRC_BYTES;
  constexpr bool STAGED = BN >= 32 && BN <= 64;  // coalesced bf16 I/O through smem staging
  constexpr int SGW = STAGED ? BN : 32;
  // A-in-TMEM mode: [2 accumulators][3 row slots][NSRC][kdim column shifts][BKC / 2 columns]
  constexpr int COPY_COLS = BKC / 2;
  constexpr int TSLOT_COLS = kdim * NSRC * COPY_COLS;
  constexpr uint32_t ATM_NEED = 2 * BN + 3 * TSLOT_COLS;
  constexpr bool ATM_OK = NTAPS == 9 && ATM_NEED <= 512;
  // TMEM row slots: 3 when the MMA warp fills them itself (tcgen05.cp executes in order with the MMAs); with
  // ELU-on-load the transform warps write the slots from registers (tcgen05.st), which needs a free / ready handshake
  // and therefore some slack
  constexpr int RT_MAX = ATM_OK ? (512 - 2 * BN) / TSLOT_COLS : 3;
  // (Round 1 also wrote the activated rows of the concat layer to TMEM from registers with tcgen05.st, 0.246 -> 0.216
  // ms; stacking the kernel rows along N (nblk) reaches 0.211 ms with the plain in-place transform: that path is gone.)
  constexpr bool TST = false;
  constexpr int RT = 3;
  constexpr uint32_t ATM_USE = 2 * BN + RT * TSLOT_COLS;
  constexpr uint32_t ATM_COLS = ATM_USE <= 128 ? 128 : ATM_USE <= 256 ? 256 : 512;
  // kernel rows as N blocks: ring of RA accumulator slots of BN columns (an output row's accumulator is complete
  // after THREE input rows; up to RA - 3 finished rows wait for the epilogue)
  constexpr bool NBLK = NTAPS == 9 && (BN == 16 || BN == 32 || BN == 64);
  constexpr int RA_N = BN == 64 ? 4 : 8;
  constexpr uint32_t NBLK_ACC_COLS = RA_N * BN;  // 128 / 256
  // transform into tensor memory: RTN operand slots of TSLOT_COLS columns behind the accumulators
  constexpr bool TSTN = ELU_IN && NBLK && BN >= 32 && NSRC == 2;  // (single-source layers: measured slower, see host side)
  constexpr int RTN_FIT = TSTN ? (512 - static_cast<int>(NBLK_ACC_COLS)) / TSLOT_COLS : 0;
  constexpr int RTN = RTN_FIT > 4 ? 4 : RTN_FIT;
  static_assert(!TSTN || RTN >= 2, "tensor-memory operand slots must at least double-buffer");
  constexpr uint32_t NBLK_COLS = TSTN ? 512u : (NBLK_ACC_COLS < 32 ? 32u : NBLK_ACC_COLS);

  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t w_full;
  __shared__ uint64_t slot_full[kRowsMaxRing];
  __shared__ uint64_t slot_empty[kRowsMaxRing];
  __shared__ uint64_t slot_ready[kRowsMaxRing];  // ELU-on-load: the slot's rows have been transformed
  __shared__ uint64_t acc_full[8];
  __shared__ uint64_t acc_empty[8];
  __shared__ uint64_t res_bar[8];  // TMA epilogue: a warp's residual tile has landed
  __shared__ uint64_t tslot_free[4];  // ELU into tensor memory: the MMAs that read an operand slot have completed
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_scale[BN];
  __shared__ float s_shift[BN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = p.ring;
  uint8_t* wsm = smem;                             // [NSRC*NTAPS][BN][ROWB]
  uint8_t* ring = smem + NSRC * NTAPS * WT_BYTES;  // [R][NSRC][SRC_BYTES]
  uint8_t* staging = ring + R * slot_bytes;        // [8 epilogue warps][2][32][BN*2+16]

  // this CTA's contiguous, balanced share of the output rows
  const long long g_begin = p.total_rows * blockIdx.x / gridDim.x;
  const long long g_end = p.total_rows * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    if (p.tma) tma_prefetch_desc(&p.tmO);
    if (p.tma > 1) tma_prefetch_desc(&p.tmR);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < R; ++i) {
      mbar_init(&slot_full[i], 1);
      mbar_init(&slot_empty[i], 1);
      mbar_init(&slot_ready[i], (TSTN && p.tst) ? 4 : 1);  // all four transform warps / the warp that owns the row
    }
    for (int i = 0; i < 4; ++i) mbar_init(&tslot_free[i], 1);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);  // one arrival per epilogue warp of the owning warpgroup
      mbar_init(&res_bar[i], 1);
    }
    fence_mbar_init();
  }
  const bool nblk = NBLK && p.nblk != 0;
  const bool tst = TSTN && nblk && p.tst != 0;
  const bool atm = ATM_OK && p.atm != 0 && !nblk;
  const int RA = nblk ? RA_N : 2;  // accumulator slots: running output row `go` lives in slot go % RA
  if (warp == 2) {
    if constexpr (NBLK) {
      if (nblk) tmem_alloc<NBLK_COLS>(&tmem_base_slot);
    }
    if (!nblk) {
      if (atm) tmem_alloc<ATM_COLS>(&tmem_base_slot);
      else tmem_alloc<TMEM_COLS>(&tmem_base_slot);
    }
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < BN; i += kRowsThreads) {
    s_scale[i] = p.ep.scale != nullptr ? p.ep.scale[i] : 1.0f;
    s_shift[i] = p.ep.shift != nullptr ? p.ep.shift[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Producer and MMA warps run their loops with ALL lanes (warp-uniform control flow); the TMA /
  // tcgen05 instructions are issued from inside `if (elect_one())` blocks.  ptxas recognises the
  // elect.sync idiom and emits back-to-back UTCHMMA with uniform-register descriptors; a plain
  // `if (lane == 0)` made it wrap every tcgen05.mma in a per-lane "waterfall" loop (~70 cycles per MMA,
  // measured: profiles/r01_exp_rates.log).
  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    constexpr int ntiles = NSRC * NTAPS;
    if (elect_one()) {
      mbar_expect_tx(&w_full, ntiles * WT_BYTES);
      if (nblk) {
        // per (source, column tap): the three kernel rows' tiles stacked along N in the order kh = 2, 1, 0 (the
        // output rows r-1, r, r+1 an input row r feeds); data gradient: geometric tap t carries weight tile 8 - t
        for (int sq = 0; sq < NSRC * 3; ++sq)
          for (int kh = 0; kh < 3; ++kh) {
            const int s = sq / 3, q = sq - 3 * s, t = kh * 3 + q;
            tma_load_2d(wsm + (sq * 3 + (2 - kh)) * WT_BYTES, &p.tmB, &w_full,
                        (s * NTAPS + (p.flip ? (NTAPS - 1) - t : t)) * BKC, 0);
          }
      } else {
        for (int i = 0; i < ntiles; ++i) tma_load_2d(wsm + i * WT_BYTES, &p.tmB, &w_full, i * BKC, 0);
      }
    }
    __syncwarp();
    constexpr uint32_t row_tx = NSRC * (kRowsW + 2 * halo) * ROWB;
    int st = 0;
    uint32_t ph = 0;
    RowWalker rw;
    for (rw.init(g_begin, g_end, p.ep.hout, p.batch); rw.valid(); rw.next_segment()) {
      const int n_in_rows = rw.seg_left + 2 * halo;
      for (int ir = 0; ir < n_in_rows; ++ir) {
        mbar_wait(&slot_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&slot_full[st], row_tx);
          uint8_t* dst = ring + st * slot_bytes;
#pragma unroll
          for (int s = 0; s < NSRC; ++s)
            tma_load_4d(dst + s * SRC_BYTES, &p.tmA[s], &slot_full[st], 0, rw.w0 - halo, rw.h - halo + ir, rw.n);
        }
        __syncwarp();
        if (++st == R) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(kRowsW, BN);
    const uint32_t desc_hi = umma_desc_hi_kmajor(ROWB);
    mbar_wait(&w_full, 0);
    const uint32_t w_lo = umma_desc_lo(smem_u32(wsm));
    const uint32_t ring_lo = umma_desc_lo(smem_u32(ring));
    constexpr uint32_t slot_units = static_cast<uint32_t>(slot_bytes) >> 4;
    int slot0 = 0;                       // ring slot of this output row's first input row
    int wait_slot = 0;                   // next slot whose arrival has not been observed yet
    uint32_t wait_ph = 0;
    int rows_seen = 0, rows_needed = 0;  // running input-row counters
    uint32_t go = 0;                     // running output-row counter
    uint64_t* const in_bar = ELU_IN ? slot_ready : slot_full;
    RowWalker rw;
    if constexpr (NBLK) {
      if (nblk) {
        // ---- kernel rows as N blocks: one pass over every input row ----
        constexpr uint32_t idesc1 = umma_idesc_bf16(kRowsW, BN), idesc2 = umma_idesc_bf16(kRowsW, 2 * BN),
                           idesc3 = umma_idesc_bf16(kRowsW, 3 * BN);
        constexpr uint32_t wt_units = static_cast<uint32_t>(WT_BYTES) >> 4;
        uint32_t go0 = 0;  // running index of the current segment's first output row
        int tt = 0;        // tensor-memory operand slot of the current input row (tst)
        for (rw.init(g_begin, g_end, p.ep.hout, p.batch); rw.valid(); rw.next_segment()) {
          const int th = rw.seg_left;
          for (int ir = 0; ir < th + 2; ++ir) {  // input row h0 - 1 + ir feeds the segment's output rows ir-2 .. ir
            mbar_wait(&in_bar[wait_slot], wait_ph);
            const int o_lo = ir >= 2 ? ir - 2 : 0;
            const int o_hi = ir < th ? ir : th - 1;
            const bool fresh = ir < th;  // output row `ir` receives its FIRST contribution here: overwrite its slot
            if (fresh) {
              const uint32_t gf = go0 + ir;
              mbar_wait(&acc_empty[gf % RA_N], ((gf / RA_N) & 1) ^ 1);
            }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_base = ring_lo + wait_slot * slot_units;
              const uint32_t a_tbase = tmem_base + NBLK_ACC_COLS + tt * TSLOT_COLS;
              // one MMA (two where the span wraps around the accumulator ring) over output rows oa .. ob; the A
              // operand comes from tensor memory (ts: a = TMEM column address) or shared memory (a = descriptor lo)
              auto span = [&](int oa, int ob, bool ts, uint32_t a, uint32_t b_lo, uint32_t acc) {
                const uint32_t sa = (go0 + oa) % RA_N;
                const int n = ob - oa + 1;
                const int j0 = oa - (ir - 2);
                const int n1 = (sa + n <= RA_N) ? n : static_cast<int>(RA_N - sa);
                const uint32_t id1 = n1 == 1 ? idesc1 : (n1 == 2 ? idesc2 : idesc3);
                if (ts) umma_bf16_ts(tmem_base + sa * BN, a, umma_desc_join(desc_hi, b_lo + j0 * wt_units), id1, acc);
                else umma_bf16(tmem_base + sa * BN, umma_desc_join(desc_hi, a), umma_desc_join(desc_hi, b_lo + j0 * wt_units), id1, acc);
                if (n1 < n) {
                  const uint32_t id2 = (n - n1) == 1 ? idesc1 : idesc2;
                  if (ts) umma_bf16_ts(tmem_base, a, umma_desc_join(desc_hi, b_lo + (j0 + n1) * wt_units), id2, acc);
                  else umma_bf16(tmem_base, umma_desc_join(desc_hi, a), umma_desc_join(desc_hi, b_lo + (j0 + n1) * wt_units), id2, acc);
                }
              };
#pragma unroll
              for (int s = 0; s < NSRC; ++s) {
                const bool ts = tst && ((p.elu_in >> s) & 1);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
#pragma unroll
                  for (int j = 0; j < BKC / 16; ++j) {
                    const uint32_t a = ts ? a_tbase + (s * 3 + q) * COPY_COLS + 8 * j
                                          : a_base + s * (SRC_BYTES >> 4) + q * (ROWB >> 4) + 2 * j;
                    const uint32_t b_lo = w_lo + (s * 3 + q) * 3 * wt_units + 2 * j;
                    if ((s | q | j) == 0) {
                      const int acc_hi = fresh ? ir - 1 : o_hi;
                      if (o_lo <= acc_hi) span(o_lo, acc_hi, ts, a, b_lo, 1u);
                      if (fresh) span(ir, ir, ts, a, b_lo, 0u);
                    } else {
                      span(o_lo, o_hi, ts, a, b_lo, 1u);
                    }
                  }
                }
              }
              if (ir >= 2) umma_commit(&acc_full[(go0 + ir - 2) % RA_N]);  // output row ir-2 is complete
              umma_commit(&slot_empty[wait_slot]);
              if (tst) umma_commit(&tslot_free[tt]);  // the operand slot may be rewritten by the transform warps
            }
            __syncwarp();
            if (++wait_slot == R) {
              wait_slot = 0;
              wait_ph ^= 1;
            }
            if (TSTN && ++tt == RTN) tt = 0;
          }
          go0 += th;
        }
      }
    }
    if constexpr (ATM_OK) {
      if (atm) {
        // ---- A operand in tensor memory ----
        // Sources that arrive already activated are copied shared -> tensor memory by this warp (tcgen05.cp, in
        // order with the MMAs).  RAW sources (ELU-on-load) are written to tensor memory by the transform warps
        // straight from registers (tcgen05.st): their rows never return to shared memory and need no copies.
        const uint32_t a_tmem0 = tmem_base + 2 * BN;
        const int elu_mask = TST ? p.elu_in : 0;
        const bool any_cp = elu_mask != (1 << NSRC) - 1;
        int t_head = 0;   // TMEM row slot of the next arriving input row
        int t_first = 0;  // TMEM row slot of the current output row's first input row
        for (rw.init(g_begin, g_end, p.ep.hout, p.batch); rw.valid(); rw.next_segment()) {
          const int th = rw.seg_left;
          for (int orow = 0; orow < th; ++orow, ++go) {
            const uint32_t b = go & 1;
            if (orow == 0) t_first = t_head;
            const int n_new = (orow == 0) ? 1 + 2 * halo : 1;
            for (int i = 0; i < n_new; ++i) {
              mbar_wait(&in_bar[wait_slot], wait_ph);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t src_lo = ring_lo + wait_slot * slot_units;
                const uint32_t dst = a_tmem0 + t_head * TSLOT_COLS;
#pragma unroll
                for (int s = 0; s < NSRC; ++s) {
                  if ((elu_mask >> s) & 1) continue;
#pragma unroll
                  for (int q = 0; q < kdim; ++q)
#pragma unroll
                    for (int j = 0; j < BKC / 16; ++j)
                      tmem_cp_128x256b(dst + (s * kdim + q) * COPY_COLS + 8 * j,
                                       umma_desc_join(desc_hi, src_lo + s * (SRC_BYTES >> 4) + q * (ROWB >> 4) + 2 * j));
                }
                // the shared-memory slot is free once the copies are done (no copies: it is free now)
                if (any_cp) umma_commit(&slot_empty[wait_slot]);
                else mbar_arrive(&slot_empty[wait_slot]);
              }
              __syncwarp();
              if (++wait_slot == R) {
                wait_slot = 0;
                wait_ph ^= 1;
              }
              if (++t_head == RT) t_head = 0;
            }
            const uint32_t d_tmem = tmem_base + b * BN;
            mbar_wait(&acc_empty[b], ((go >> 1) & 1) ^ 1);
            tc_fence_after();
            int ts[3];
            ts[0] = t_first;
            ts[1] = t_first + 1 >= RT ? t_first + 1 - RT : t_first + 1;
            ts[2] = t_first + 2 >= RT ? t_first + 2 - RT : t_first + 2;
            if (elect_one()) {
#pragma unroll
              for (int s = 0; s < NSRC; ++s) {
#pragma unroll
                for (int r = 0; r < kdim; ++r) {
                  const uint32_t a_row = a_tmem0 + ts[r] * TSLOT_COLS + s * kdim * COPY_COLS;
#pragma unroll
                  for (int q = 0; q < kdim; ++q) {
                    const int wt = p.flip ? (NTAPS - 1) - (r * kdim + q) : r * kdim + q;
                    const uint32_t b_lo = w_lo + (s * NTAPS + wt) * (WT_BYTES >> 4);
#pragma unroll
                    for (int j = 0; j < BKC / 16; ++j)
                      umma_bf16_ts(d_tmem, a_row + q * COPY_COLS + 8 * j, umma_desc_join(desc_hi, b_lo + 2 * j), idesc,
                                   (s | r | q | j) != 0 ? 1u : 0u);
                  }
                }
              }
              umma_commit(&acc_full[b]);
            }
            __syncwarp();
            t_first = ts[1];
          }
        }
      }
    }
    if (!atm && !nblk)
    for (rw.init(g_begin, g_end, p.ep.hout, p.batch); rw.valid(); rw.next_segment()) {
      const int th = rw.seg_left;
      for (int orow = 0; orow < th; ++orow, ++go) {
        const uint32_t b = go & 1;
        mbar_wait(&acc_empty[b], ((go >> 1) & 1) ^ 1);
        rows_needed += (orow == 0) ? 1 + 2 * halo : 1;
        for (; rows_seen < rows_needed; ++rows_seen) {
          mbar_wait(&in_bar[wait_slot], wait_ph);
          if (++wait_slot == R) {
            wait_slot = 0;
            wait_ph ^= 1;
          }
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * BN;
        // ring slots of the kdim input rows this output row reads
        int sl[3];
        sl[0] = slot0;
        sl[1] = slot0 + 1 >= R ? slot0 + 1 - R : slot0 + 1;
        sl[2] = slot0 + 2 >= R ? slot0 + 2 - R : slot0 + 2;
        const bool last = orow == th - 1;
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < NSRC; ++s) {
#pragma unroll
            for (int r = 0; r < kdim; ++r) {
              const uint32_t a_lo = ring_lo + sl[r] * slot_units + s * (SRC_BYTES >> 4);
#pragma unroll
              for (int q = 0; q < kdim; ++q) {
                // data gradient: the geometric tap (r, q) carries the weights of the mirrored tap
                const int wt = p.flip ? (NTAPS - 1) - (r * kdim + q) : r * kdim + q;
                const uint32_t b_lo = w_lo + (s * NTAPS + wt) * (WT_BYTES >> 4);
#pragma unroll
                for (int j = 0; j < BKC / 16; ++j)
                  umma_bf16(d_tmem, umma_desc_join(desc_hi, a_lo + q * (ROWB >> 4) + 2 * j),
                            umma_desc_join(desc_hi, b_lo + 2 * j), idesc, (s | r | q | j) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&acc_full[b]);
          umma_commit(&slot_empty[sl[0]]);  // input row `orow` is not needed by later output rows
          if (last && halo) {               // the segment's last output row also releases the halo rows
            umma_commit(&slot_empty[sl[1]]);
            umma_commit(&slot_empty[sl[2]]);
          }
        }
        __syncwarp();
        slot0 = sl[1];
        if (last && halo) slot0 = sl[2] + 1 >= R ? sl[2] + 1 - R : sl[2] + 1;
      }
    }
  } else if (warp >= 10) {
    // ------------------------------ ELU-on-load transform ---------------------------------------
    // Elementwise and in place, so the TMA swizzle inside the slot does not matter: the box footprint is the
    // contiguous first (128 + 2*halo) * ROWB bytes of each source's buffer.
    if constexpr (!ELU_IN) return;
    if constexpr (TSTN) {
      if (tst) {
        // ---- transform from shared memory straight INTO tensor memory (see RowsKParams::tst) ----
        // TMEM lane m of column-tap copy q holds box pixel m + q.  Each warp serves its own lane quarter: a lane
        // loads and activates box pixel 32k + lane (and, one word per lane, the two pixels past the quarter), the
        // shifted copies come from warp shuffles, and three tcgen05.st per 32 channels write the copies.
        const int k = warp & 3;
        const uint32_t a_tmem0 = tmem_base + NBLK_ACC_COLS + (static_cast<uint32_t>(k * 32) << 16);
        constexpr int GROUPS = ROWB / 64;  // groups of 32 channels = 16 packed registers = 4 sixteen-byte chunks
        int st = 0, tt = 0;
        uint32_t ph = 0, tph = 0;
        RowWalker rw;
        for (rw.init(g_begin, g_end, p.ep.hout, p.batch); rw.valid(); rw.next_segment()) {
          const int n_in_rows = rw.seg_left + 2 * halo;
          for (int ir = 0; ir < n_in_rows; ++ir) {
            mbar_wait(&slot_full[st], ph);
            mbar_wait(&tslot_free[tt], tph ^ 1);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < NSRC; ++s) {
              if (!((p.elu_in >> s) & 1)) continue;
              const uint32_t base = smem_u32(ring + st * slot_bytes + s * SRC_BYTES);
              const int j0 = k * 32 + lane;                              // own box pixel
              const int key0 = ROWB == 64 ? ((j0 >> 1) & 3) : (j0 & 7);   // TMA swizzle of the pixel's row
              // the two pixels past the quarter (box pixels 32k+32, 32k+33): 2 x 16 packed words per channel group,
              // ONE word per lane (lane = 16 * pixel + word) instead of a second full pass over 32 lanes
              const int j1 = k * 32 + 32 + (lane >> 4);
              const int key1 = ROWB == 64 ? ((j1 >> 1) & 3) : (j1 & 7);
#pragma unroll
              for (int g = 0; g < GROUPS; ++g) {
                uint32_t e[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  uint4 a = lds128(base + j0 * ROWB + (((4 * g + c) ^ key0) << 4));
                  elu_bf16x8(a);
                  e[4 * c] = a.x; e[4 * c + 1] = a.y; e[4 * c + 2] = a.z; e[4 * c + 3] = a.w;
                }
                uint32_t hv;
                {
                  const int w = 16 * g + (lane & 15);  // 32-bit word (2 channels) of the halo pixel
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hv)
                               : "r"(base + j1 * ROWB + ((((w >> 2) ^ key1) << 4) | ((w & 3) << 2))) : "memory");
                  float f0, f1;
                  unpack_bf16x2(hv, f0, f1);
                  hv = pack_bf16x2(elu1(f0), elu1(f1));
                }
                uint32_t v1[16], v2[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint32_t x1 = __shfl_down_sync(0xffffffffu, e[i], 1);
                  const uint32_t x2 = __shfl_down_sync(0xffffffffu, e[i], 2);
                  const uint32_t b0 = __shfl_sync(0xffffffffu, hv, i);        // halo pixel 0, word i
                  const uint32_t b1 = __shfl_sync(0xffffffffu, hv, 16 + i);   // halo pixel 1, word i
                  v1[i] = lane == 31 ? b0 : x1;
                  v2[i] = lane == 30 ? b0 : (lane == 31 ? b1 : x2);
                }
                const uint32_t dst = a_tmem0 + tt * TSLOT_COLS + s * kdim * COPY_COLS + 16 * g;
                tmem_st16(dst, e);
                tmem_st16(dst + COPY_COLS, v1);
                tmem_st16(dst + 2 * COPY_COLS, v2);
              }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_ready[st]);
            if (++st == R) {
              st = 0;
              ph ^= 1;
            }
            if (++tt == RTN) {
              tt = 0;
              tph ^= 1;
            }
          }
        }
      }
    }
    if (!tst) {
    // Row per warp: warp tw transforms the input rows tw, tw + 4, ... on its own (16-byte chunks, 4 in flight per
    // lane).  Splitting every row over the four warps instead made each of them pay the per-row overhead
    // (barrier wait, fence.proxy.async, arrive).  A row costs one warp ~2800 cycles (measured: with TWO transform
    // warps the 1x1 layer fell from 945 to 1400 cycles per row), so four warps are needed for a ~900-cycle row period.
    const int tw = warp - 10;
    constexpr int n_chunks = (kRowsW + 2 * halo) * ROWB / 16;
    int st = 0;
    uint32_t ph = 0;
    int rowc = 0;  // running input-row counter
    RowWalker rw;
    for (rw.init(g_begin, g_end, p.ep.hout, p.batch); rw.valid(); rw.next_segment()) {
      const int n_in_rows = rw.seg_left + 2 * halo;
      for (int ir = 0; ir < n_in_rows; ++ir, ++rowc) {
        if ((rowc & 3) == tw) {
          mbar_wait(&slot_full[st], ph);
#pragma unroll
          for (int s = 0; s < NSRC; ++s) {
            if (!((p.elu_in >> s) & 1)) continue;
            const uint32_t base = smem_u32(ring + st * slot_bytes + s * SRC_BYTES);
            // 8 independent 16-byte chunks (64 elements) in flight per lane: the loop is latency-bound (ncu: the
            // transform warps issue at 0.18 IPC, stalled on the congested shared-memory pipe and the MUFU results)
            constexpr int UNR = 8;
            constexpr int n_full = n_chunks / (32 * UNR);
#pragma unroll 1
            for (int it = 0; it < n_full; ++it) {
              const uint32_t a0 = base + 16 * (it * 32 * UNR + lane);
              uint4 v[UNR];
#pragma unroll
              for (int u = 0; u < UNR; ++u) v[u] = lds128(a0 + u * 512);
#pragma unroll
              for (int u = 0; u < UNR; ++u) elu_bf16x8(v[u]);
#pragma unroll
              for (int u = 0; u < UNR; ++u) sts128(a0 + u * 512, v[u]);
            }
            constexpr int n_tail = n_chunks - n_full * 32 * UNR;  // the halo pixels: 8 (C = 32) or 16 (C = 64) chunks
            static_assert(n_tail <= 32, "tail must fit one pass");
            if (n_tail > 0 && lane < n_tail) {
              const uint32_t a0 = base + 16 * (n_full * 32 * UNR + lane);
              uint4 v = lds128(a0);
              elu_bf16x8(v);
              sts128(a0, v);
            }
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&slot_ready[st]);
        }
        if (++st == R) {
          st = 0;
          ph ^= 1;
        }
      }
    }
    }  // in-place transform (operand stays in shared memory)
  } else {
    // ------------------------------ epilogue (two warpgroups, alternating rows) ----------------
    const int g = (warp - 2) >> 2;  // warpgroup 0: warps 2-5 (even running rows), 1: warps 6-9 (odd)
    const int q = warp & 3;         // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);  // + slot * BN: this row's accumulator
    const bool has_res = p.ep.residual != nullptr;
    const bool bf16_out = (p.ep.flags & (VUB_CONV_F32_NHWC | VUB_CONV_F32_NCHW | VUB_CONV_U8_HWC)) == 0;
    const uint32_t t0 = smem_u32(staging) + (warp - 2) * EpiStage<SGW>::WARP_BYTES;  // residual in / raw out
    const uint32_t t1 = t0 + EpiStage<SGW>::TILE_BYTES;                               // elu out
    constexpr int CPR = SGW / 8;
    uint4 rblk[CPR];  // STAGED: this warp's 32-pixel residual block of its NEXT row, linear 16B chunks

    // element offset of channel 0 of pixel (n, oh, w0 + q*32) == start of this warp's contiguous block
    auto block_off = [&](const RowWalker& r) {
      return (((long long)r.n * p.ep.hout + r.h) * p.ep.wout + r.w0 + q * 32) * (long long)p.ep.cout;
    };
    auto prefetch_res = [&](const RowWalker& r) {
      const uint4* src = reinterpret_cast<const uint4*>(p.ep.residual + block_off(r));
#pragma unroll
      for (int k = 0; k < CPR; ++k) rblk[k] = __ldg(src + lane + 32 * k);
    };

    RowWalker cur, ahead;  // `ahead` runs two rows in front (this warpgroup's next row)
    cur.init(g_begin, g_end, p.ep.hout, p.batch);
    if (g == 1 && cur.valid()) cur.next_row();
    if constexpr (BWD && (BN == 32 || BN == 64)) {
      // ---- data-gradient mode: dx = ELU'(mask) * acc + addend + addend2 per source (half) ----
      // The warp's 32 pixels are ONE contiguous block of every tensor, so all addresses are `block + lane * C`.
      // The epilogue inputs do not depend on the accumulator: the first 32-channel group's loads are issued
      // BEFORE waiting for the MMA, the next group's as soon as the registers are free again.
      const uint32_t st = smem_u32(staging) + (warp - 2) * p.stage_tiles * EpiStage<32>::TILE_BYTES;
      const bool two = p.n_halves == 2;
      constexpr int NG = BN / 32;
      uint32_t gob = g;  // running index of this warpgroup's current row (g, g + 2, ...)
      uint4 va[4], vm[4], vb[4];
      while (cur.valid()) {
        const uint32_t aslot = gob % RA;
        const uint32_t trow = tlane + aslot * BN;
        const long long pix = ((long long)cur.n * p.ep.hout + cur.h) * p.ep.wout + cur.w0 + q * 32;  // warp's first pixel
        auto load_group = [&](int gi) {
          const TileEpi& e = p.bwd[two ? gi : 0];
          const long long off = (pix + lane) * e.out_c + (two ? 0 : gi * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            va[j] = vm[j] = vb[j] = make_uint4(0u, 0u, 0u, 0u);
            if (e.addend != nullptr) va[j] = *reinterpret_cast<const uint4*>(e.addend + off + 8 * j);  // may alias dx
            if (e.mask_elu != nullptr) vm[j] = __ldg(reinterpret_cast<const uint4*>(e.mask_elu + off + 8 * j));
            if (e.addend2 != nullptr) vb[j] = *reinterpret_cast<const uint4*>(e.addend2 + off + 8 * j);
          }
        };
        load_group(0);
        mbar_wait(&acc_full[aslot], (gob / RA) & 1);
        tc_fence_after();
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          const TileEpi& e = p.bwd[two ? gi : 0];
          uint32_t acc[32];
          tmem_ld32(trow + gi * 32, acc);
          tmem_ld_wait();
          if (gi == NG - 1) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[aslot]);
          }
          float y[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(acc[i]);
          if (e.mask_elu != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float ev[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              add_bf16x8(vm[j], ev);
#pragma unroll
              for (int i = 0; i < 8; ++i) y[8 * j + i] *= act1_grad(ev[i], (e.flags & VUB_CONV_RELU) != 0);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            add_bf16x8(va[j], &y[8 * j]);
            add_bf16x8(vb[j], &y[8 * j]);
          }
          if (gi + 1 < NG) load_group(gi + 1);
          if (two || NG == 1) {  // a 32-channel half (or the whole 32-channel tensor): stage, flush
            stage_put32<32>(st, lane, 0, y);
            __syncwarp();
            stage_flush_linear<32>(st, e.out_raw + pix * 32, lane);
            __syncwarp();
          } else {               // one 64-channel tensor in two groups
            stage_put32<64>(st, lane, gi * 32, y);
            if (gi == NG - 1) {
              __syncwarp();
              stage_flush_linear<64>(st, e.out_raw + pix * 64, lane);
              __syncwarp();
            }
          }
        }
        gob += 2;
        cur.next_row();
        if (cur.valid()) cur.next_row();
      }
    }
    if constexpr (!BWD) {
    ahead = cur;
    // ---- TMA epilogue state ----
    constexpr bool TMA_OK = STAGED;  // BN 32 / 64
    const bool one_out = (p.ep.out_raw != nullptr) != (p.ep.out_elu != nullptr);
    const int tma = (TMA_OK && bf16_out && one_out) ? p.tma : 0;
    const bool tma_res = tma > 1 && has_res;
    const int wi = warp - 2;
    const uint32_t tR = t0, tO = has_res ? t1 : t0;  // TMA mode: residual tile / output tile
    uint32_t rph = 0;
    // box coordinate of a warp's 32-pixel block in the [pixels (pairs at BN = 32)] x [128 bytes] view
    auto box_row = [&](const RowWalker& r) {
      const long long pix = ((long long)r.n * p.ep.hout + r.h) * p.ep.wout + r.w0 + q * 32;
      return static_cast<int>(BN == 32 ? pix >> 1 : pix);
    };
    auto tma_res_issue = [&](const RowWalker& r) {
      if (lane == 0) {
        mbar_expect_tx(&res_bar[wi], TmaStage<SGW>::TILE_BYTES);
        tma_load_2d(reinterpret_cast<void*>(smem + (tR - smem_u32(smem))), &p.tmR, &res_bar[wi], 0, box_row(r));
      }
    };
    // (Per-channel affine in registers instead of broadcast reads of s_scale / s_shift was measured SLOWER: 64 more live
    // registers per lane cost the 320-thread kernels spills and ILP, 32->32 pair I/O 0.174 -> 0.187 ms.)
    constexpr bool AFF_REG = false;
    float r_scale[AFF_REG ? BN : 1], r_shift[AFF_REG ? BN : 1];
    if constexpr (AFF_REG) {
#pragma unroll
      for (int i = 0; i < BN; ++i) {
        r_scale[i] = s_scale[i];
        r_shift[i] = s_shift[i];
      }
    }
    if (STAGED && has_res && bf16_out && cur.valid()) {
      if (tma_res) tma_res_issue(cur);
      else prefetch_res(cur);
    }
    uint32_t gof = g;  // running index of this warpgroup's current row (g, g + 2, ...)
    while (cur.valid()) {
      const uint32_t aslot = gof % RA;
      const uint32_t trow = tlane + aslot * BN;
      if (ahead.valid()) ahead.next_row();
      if (ahead.valid()) ahead.next_row();
      const int n = cur.n, oh = cur.h, ow = cur.w0 + m;
      // (Loading the residual straight into per-lane registers instead of staging it was measured: no gain at
      // BN = 32 and 18 % slower at BN = 64 — 32 cache lines per load instruction; profiles/r01_ncu_rows_smem.txt.)
      if constexpr (STAGED) {
        if (has_res && bf16_out) {
          if (tma_res) {
            mbar_wait(&res_bar[wi], rph);
            rph ^= 1;
          } else if (tma) {
#pragma unroll
            for (int k = 0; k < CPR; ++k) {  // linear 16-byte chunks of the block -> TMA-layout tile
              const int cidx = lane + 32 * k;
              sts128(tR + TmaStage<SGW>::off(cidx / CPR, cidx % CPR), rblk[k]);
            }
            __syncwarp();
            if (ahead.valid()) prefetch_res(ahead);
          } else {
            stage_fill_linear<SGW>(t0, rblk, lane);  // residual block -> staging; lanes then read their own row
            __syncwarp();
            if (ahead.valid()) prefetch_res(ahead);
          }
        }
      }
      mbar_wait(&acc_full[aslot], (gof / RA) & 1);
      tc_fence_after();
      if constexpr (STAGED) {
        if (tma) {
          // the previous tile's store must have finished reading the output tile before it is rewritten
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
      }
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        uint32_t acc[CH];
        if constexpr (CH == 32) tmem_ld32(trow + c * CH, acc);
        else tmem_ld16(trow + c * CH, acc);
        tmem_ld_wait();
        if (c == NCHUNK - 1) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[aslot]);
        }
        float y[CH];
        if constexpr (AFF_REG) {
#pragma unroll
          for (int i = 0; i < CH; ++i) y[i] = fmaf(__uint_as_float(acc[i]), r_scale[c * CH + i], r_shift[c * CH + i]);
        } else {
#pragma unroll
          for (int i = 0; i < CH; ++i) y[i] = fmaf(__uint_as_float(acc[i]), s_scale[c * CH + i], s_shift[c * CH + i]);
        }
        if constexpr (STAGED) {
          if (bf16_out) {
            if (tma) {
              if (has_res) {
#pragma unroll
                for (int j = 0; j < CH / 8; ++j)
                  add_bf16x8(lds128(tR + TmaStage<SGW>::off(lane, c * (CH / 8) + j)), &y[8 * j]);
              }
              if (p.ep.out_elu != nullptr) {
                const bool relu = (p.ep.flags & VUB_CONV_RELU) != 0;
#pragma unroll
                for (int i = 0; i < CH; ++i) y[i] = act1(y[i], relu);
              }
              stage_put32_tma<SGW>(tO, lane, c * CH, y);
              continue;
            }
            if (has_res) {
#pragma unroll
              for (int j = 0; j < CH / 8; ++j)
                add_bf16x8(lds128(t0 + EpiStage<SGW>::off(lane, c * (CH / 8) + j)), &y[8 * j]);
            }
            if (p.ep.out_raw != nullptr) stage_put32<SGW>(t0, lane, c * CH, y);
            if (p.ep.out_elu != nullptr) {
              const bool relu = (p.ep.flags & VUB_CONV_RELU) != 0;
#pragma unroll
              for (int i = 0; i < CH; ++i) y[i] = act1(y[i], relu);
              stage_put32<SGW>(t1, lane, c * CH, y);
            }
            continue;
          }
        }
        if (has_res) {
          const long long pix = ((long long)n * p.ep.hout + oh) * p.ep.wout + ow;
          const uint4* r = reinterpret_cast<const uint4*>(p.ep.residual + pix * p.ep.cout + c * CH);
#pragma unroll
          for (int j = 0; j < CH / 8; ++j) add_bf16x8(__ldg(r + j), &y[8 * j]);
        }
        epi_store<CH>(p.ep, y, c * CH, n, oh, ow);
      }
      if constexpr (STAGED) {
        if (bf16_out) {
          if (tma) {
            fence_proxy_async();  // this lane's tile writes -> visible to the TMA engine
            __syncwarp();         // ... all 32 lanes' (and every lane has read its residual chunks)
            if (lane == 0) {
              tma_store_2d(&p.tmO, tO, 0, box_row(cur));
              bulk_commit();
            }
            if (tma_res && ahead.valid()) tma_res_issue(ahead);  // the residual tile is free again
          } else {
            const long long off = block_off(cur);
            __syncwarp();
            if (p.ep.out_raw != nullptr) stage_flush_linear<SGW>(t0, p.ep.out_raw + off, lane);
            if (p.ep.out_elu != nullptr) stage_flush_linear<SGW>(t1, p.ep.out_elu + off, lane);
            __syncwarp();
          }
        }
      }
      gof += 2;
      cur = ahead;
    }
    if constexpr (STAGED) {
      if (tma && lane == 0) bulk_wait<0>();  // all tiles written before the CTA's shared memory goes away
    }
    }  // !BWD
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    if constexpr (NBLK) {
      if (nblk) tmem_dealloc<NBLK_COLS>(tmem_base);
    }
    if (!nblk) {
      if (atm) tmem_dealloc<ATM_COLS>(tmem_base);
      else tmem_dealloc<TMEM_COLS>(tmem_base);
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int BN, int BKC, int NTAPS, int NSRC, int MODE = kRowsFwd>
static int launch_rows_t(const RowsKParams& kp, int grid, size_t smem, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_rows_kernel<BN, BKC, NTAPS, NSRC, MODE>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "cudaFuncSetAttribute(conv_rows<%d,%d>): %s", BN, BKC,
                cudaGetErrorString(attr_err));
  conv_rows_kernel<BN, BKC, NTAPS, NSRC, MODE><<<grid, MODE == kRowsFwdElu ? kRowsThreadsElu : kRowsThreads, smem, stream>>>(kp);
  return check_launch("conv_rows");
}

template <int BN, int BKC>
static int launch_rows(const RowsKParams& kp, int grid, size_t smem, cudaStream_t stream) {
  if (kp.n_halves > 0) {  // data gradient: one source (dy), 32 or 64 accumulator columns
    if constexpr (BN == 32 || BN == 64) {
      if (kp.nsrc == 1 && kp.ntaps == 9) return launch_rows_t<BN, BKC, 9, 1, kRowsBwd>(kp, grid, smem, stream);
      if (kp.nsrc == 1 && kp.ntaps == 1) return launch_rows_t<BN, BKC, 1, 1, kRowsBwd>(kp, grid, smem, stream);
    }
    return fail(VUB_E_SHAPE, "conv_rows: no data-gradient variant for BN %d BKC %d taps %d nsrc %d", BN, BKC, kp.ntaps,
                kp.nsrc);
  }
  if (kp.elu_in) {  // ELU-on-load: the square layers of the residual blocks and their 1x1 skip convs
    if constexpr (BN == BKC && (BN == 32 || BN == 64)) {
      if (kp.ntaps == 9 && kp.nsrc == 1) return launch_rows_t<BN, BKC, 9, 1, kRowsFwdElu>(kp, grid, smem, stream);
      if (kp.ntaps == 1 && kp.nsrc == 1) return launch_rows_t<BN, BKC, 1, 1, kRowsFwdElu>(kp, grid, smem, stream);
      if constexpr (BN == 32) {
        if (kp.ntaps == 9 && kp.nsrc == 2) return launch_rows_t<BN, BKC, 9, 2, kRowsFwdElu>(kp, grid, smem, stream);
      }
    }
    return fail(VUB_E_FLAGS, "conv_rows: no ELU-on-load variant for BN %d BKC %d taps %d nsrc %d", BN, BKC, kp.ntaps,
                kp.nsrc);
  }
  if (kp.ntaps == 9 && kp.nsrc == 1) return launch_rows_t<BN, BKC, 9, 1>(kp, grid, smem, stream);
  if (kp.ntaps == 9 && kp.nsrc == 2) return launch_rows_t<BN, BKC, 9, 2>(kp, grid, smem, stream);
  if (kp.ntaps == 1 && kp.nsrc == 1) return launch_rows_t<BN, BKC, 1, 1>(kp, grid, smem, stream);
  return launch_rows_t<BN, BKC, 1, 2>(kp, grid, smem, stream);
}


static size_t rows_staging_bytes(int bn) {
  return (bn >= 32 && bn <= 64) ? (size_t)8 * 2 * 32 * (bn * 2) : 0;
}

// One launch of the row-streaming kernel, forward or data-gradient mode.
struct RowsJob {
  const void* x[2];     // A sources, NHWC bf16 [batch, h, w, c_src]
  int nsrc, c_src;
  int batch, h, w;
  const void* w_packed; // [bn rows][nsrc * taps * c_src] bf16
  int bn, taps;
  EpiParams ep;         // forward epilogue (scale / shift always come from here; may be NULL = identity)
  int flip, n_halves, stage_tiles;
  TileEpi bwd[2];
  int elu_in;           // VUB_CONV_ELU_IN bits, as a source mask
};

static bool rows_shape_ok(int nsrc, int c_src, int bn, int taps, int h, int w, size_t staging) {
  if (env_int("VUB_DISABLE_ROWS", 0)) return false;
  if (w < kRowsW || h < 2 || c_src > 64 || bn > 128) return false;
  if (c_src == 32 && bn > 64) return false;
  const size_t wbytes = (size_t)nsrc * taps * bn * c_src * 2;
  const size_t slot = (size_t)nsrc * (c_src == 64 ? 136 * 128 : 144 * 64);
  const int min_ring = taps == 9 ? 4 : 2;
  return wbytes + min_ring * slot + staging + 2048 <= 224 * 1024;
}

static int rows_job_launch(const RowsJob& j, cudaStream_t stream) {
  const int bkc = j.c_src;
  RowsKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.ep = j.ep;
  kp.ep.hout = j.h;
  kp.ep.wout = j.w;
  kp.batch = j.batch;
  kp.nsrc = j.nsrc;
  kp.ntaps = j.taps;
  kp.strips_w = j.w / kRowsW;
  kp.total_rows = (long long)kp.strips_w * j.batch * j.h;
  kp.flip = j.flip;
  kp.n_halves = j.n_halves;
  kp.stage_tiles = j.stage_tiles;
  kp.bwd[0] = j.bwd[0];
  kp.bwd[1] = j.bwd[1];
  kp.elu_in = j.elu_in & ((1 << j.nsrc) - 1);
  kp.nblk = (j.taps == 9 && (j.bn == 16 || j.bn == 32 || j.bn == 64) && env_int("VUB_ROWS_NBLK", 1)) ? 1 : 0;
  // measured (B=64): into-tensor-memory wins for the concat layer (one raw + one activated source: 0.241 -> 0.222 ms)
  // and loses for the single-source layers (32->32 0.168 -> 0.175, 64->64 0.084 -> 0.111 ms: all four warps serialise
  // on one row's load -> ELU -> shuffle -> store chain), so only the two-source instantiation carries it
  kp.tst = (kp.nblk && kp.elu_in && j.bn >= 32 && j.nsrc == 2 && env_int("VUB_ROWS_TST", 1)) ? 1 : 0;
  kp.atm = (!kp.nblk && j.taps == 9 && 2 * j.bn + 3 * 3 * j.nsrc * (bkc / 2) <= 512 && env_int("VUB_ROWS_ATM", 1)) ? 1 : 0;
  const size_t wbytes = (size_t)j.nsrc * j.taps * j.bn * bkc * 2;
  const size_t slot = (size_t)j.nsrc * (bkc == 64 ? 136 * 128 : 144 * 64);
  const int ring_env = env_int("VUB_ROWS_RING", 0);
  size_t staging = rows_staging_bytes(j.bn);
  if (j.n_halves > 0) staging = (size_t)8 * j.stage_tiles * EpiStage<32>::TILE_BYTES;
  int ring = static_cast<int>((216 * 1024 - wbytes - staging) / slot);
  if (ring_env > 0) ring = ring_env;
  const int min_ring = j.taps == 9 ? 4 : 2;
  if (ring > kRowsMaxRing) ring = kRowsMaxRing;
  if (ring < min_ring) ring = min_ring;
  kp.ring = ring;
  const size_t smem = wbytes + ring * slot + staging + 1024;

  const uint64_t C = j.c_src, Hin = j.h, Win = j.w, B = j.batch;
  const uint64_t dims[4] = {C, Win, Hin, B};
  const uint64_t strides[3] = {C * 2, Win * C * 2, Hin * Win * C * 2};
  const uint32_t boxA[4] = {(uint32_t)bkc, (uint32_t)(kRowsW + (j.taps == 9 ? 2 : 0)), 1u, 1u};
  int rc;
  for (int s = 0; s < j.nsrc; ++s) {
    rc = encode_tmap_bf16(&kp.tmA[s], j.x[s], 4, dims, strides, boxA, bkc * 2, false);
    if (rc) return rc;
  }
  const uint64_t ktot = (uint64_t)j.nsrc * j.taps * bkc;
  const uint64_t dimsb[2] = {ktot, (uint64_t)j.bn};
  const uint64_t stridesb[1] = {ktot * 2};
  const uint32_t boxB[2] = {(uint32_t)bkc, (uint32_t)j.bn};
  rc = encode_tmap_bf16(&kp.tmB, j.w_packed, 2, dimsb, stridesb, boxB, bkc * 2, true);
  if (rc) return rc;

  // TMA epilogue: forward, exactly one bf16 NHWC output of 32 / 64 channels (the inference path's lean activations)
  kp.tma = 0;
  const bool bf16_one_out = j.n_halves == 0 && (j.ep.flags & (VUB_CONV_F32_NHWC | VUB_CONV_F32_NCHW | VUB_CONV_U8_HWC |
                                                               VUB_CONV_D2S)) == 0 &&
                            ((j.ep.out_raw != nullptr) != (j.ep.out_elu != nullptr));
  if (bf16_one_out && (j.bn == 32 || j.bn == 64)) {
    kp.tma = env_int("VUB_ROWS_TMA", 2);
    if (kp.tma) {
      const uint64_t total_pix = (uint64_t)j.batch * j.h * j.w;
      const uint64_t rows128 = j.bn == 32 ? total_pix / 2 : total_pix;  // 128-byte rows: pixel pairs at 32 channels
      const uint64_t dimso[2] = {64, rows128};
      const uint64_t strideso[1] = {128};
      const uint32_t boxo[2] = {64, (uint32_t)(j.bn == 32 ? 16 : 32)};
      void* outp = j.ep.out_raw != nullptr ? (void*)j.ep.out_raw : (void*)j.ep.out_elu;
      rc = encode_tmap_bf16(&kp.tmO, outp, 2, dimso, strideso, boxo, 128, false);
      if (rc) return rc;
      if (kp.tma > 1 && j.ep.residual != nullptr) {
        rc = encode_tmap_bf16(&kp.tmR, j.ep.residual, 2, dimso, strideso, boxo, 128, false);
        if (rc) return rc;
      }
    }
  }

  const int n_sm = sm_count();
  // persistent: one CTA per SM, each with a contiguous balanced range of output rows (>= 2 rows each)
  const int grid_env = env_int("VUB_ROWS_GRID", 0);
  long long want = kp.total_rows / 2 < n_sm ? kp.total_rows / 2 : n_sm;
  if (grid_env > 0) want = grid_env;
  const int grid = want < 1 ? 1 : static_cast<int>(want);
  const int bn = j.bn;
#define VUB_ROWS_DISPATCH(BN_, BKC_) \
  if (bn == BN_ && bkc == BKC_) return launch_rows<BN_, BKC_>(kp, grid, smem, stream);
  VUB_ROWS_DISPATCH(16, 32)
  VUB_ROWS_DISPATCH(32, 32)
  VUB_ROWS_DISPATCH(64, 32)
  VUB_ROWS_DISPATCH(16, 64)
  VUB_ROWS_DISPATCH(32, 64)
  VUB_ROWS_DISPATCH(64, 64)
  VUB_ROWS_DISPATCH(128, 64)
#undef VUB_ROWS_DISPATCH
  return fail(VUB_E_SHAPE, "conv_rows: no kernel for BN %d BKC %d", bn, bkc);
}

// Returns true when the row-streaming kernel covers this (already validated) descriptor.
bool conv_rows_applicable(const vub_conv_desc* d) {
  if (d->stages < 0 || d->stride != 1) return false;
  if (d->flags & (VUB_CONV_ELU_IN0 | VUB_CONV_ELU_IN1)) {
    // variants instantiated in launch_rows
    const bool square = d->cout == d->c_src && (d->cout == 32 || d->cout == 64);
    const bool shape = (d->nsrc == 1 && (d->ksize == 3 || d->ksize == 1)) || (d->nsrc == 2 && d->ksize == 3 && d->cout == 32);
    if (!square || !shape) return false;
  }
  return rows_shape_ok(d->nsrc, d->c_src, d->cout, d->ksize * d->ksize, d->hin, d->win, rows_staging_bytes(d->cout));
}

int conv_rows_launch(const vub_conv_desc* d, cudaStream_t stream) {
  RowsJob j;
  memset(&j, 0, sizeof(j));
  j.x[0] = d->x[0];
  j.x[1] = d->x[1];
  j.nsrc = d->nsrc;
  j.c_src = d->c_src;
  j.batch = d->batch;
  j.h = d->hin;
  j.w = d->win;
  j.w_packed = d->w;
  j.bn = d->cout;
  j.taps = d->ksize * d->ksize;
  j.ep.scale = d->scale;
  j.ep.shift = d->shift;
  j.ep.residual = static_cast<const __nv_bfloat16*>(d->residual);
  j.ep.out_raw = static_cast<__nv_bfloat16*>(d->out_raw);
  j.ep.out_elu = static_cast<__nv_bfloat16*>(d->out_elu);
  j.ep.out_f32 = static_cast<float*>(d->out_f32);
  j.ep.cout = d->cout;
  j.ep.cout_valid = d->cout_valid;
  j.ep.flags = d->flags;
  j.elu_in = ((d->flags & VUB_CONV_ELU_IN0) ? 1 : 0) | ((d->flags & VUB_CONV_ELU_IN1) ? 2 : 0);
  return rows_job_launch(j, stream);
}

// Data gradient of a stride-1 conv on the wide, thin layers (dy has 32 or 64 channels, width >= 128):
// returns true and sets *rc when the row-streaming kernel took the job (one launch per <= 64 output columns).
bool conv_rows_dgrad(const vub_dgrad_desc* d, cudaStream_t stream, int* rc) {
  const int C = d->c_src, taps = d->ksize * d->ksize;
  if (d->stride != 1 || (d->flags & VUB_CONV_D2S) || (d->cout_k != 32 && d->cout_k != 64) || (C != 32 && C != 64))
    return false;
  bool any_addend2 = false;
  for (int s = 0; s < d->nsrc; ++s) any_addend2 = any_addend2 || d->addend2[s] != nullptr;
  const int stage_tiles = any_addend2 ? 3 : 2;
  const size_t staging = (size_t)8 * stage_tiles * EpiStage<32>::TILE_BYTES;
  const bool merged = d->nsrc * C <= 64;  // both halves of a concat in one launch (N = 64)
  const int bn = merged ? d->nsrc * C : C;
  if (!rows_shape_ok(1, d->cout_k, bn, taps, d->hin, d->win, staging)) return false;
  const int n_launch = merged ? 1 : d->nsrc;
  for (int l = 0; l < n_launch; ++l) {
    RowsJob j;
    memset(&j, 0, sizeof(j));
    j.x[0] = d->dy;
    j.nsrc = 1;
    j.c_src = d->cout_k;
    j.batch = d->batch;
    j.h = d->hin;
    j.w = d->win;
    j.w_packed = static_cast<const char*>(d->wd) + (size_t)l * C * taps * d->cout_k * 2;  // rows of source l
    j.bn = bn;
    j.taps = taps;
    j.ep.cout = bn;
    j.ep.cout_valid = bn;
    j.flip = 1;
    j.n_halves = merged ? d->nsrc : 1;
    j.stage_tiles = stage_tiles;
    for (int hh = 0; hh < j.n_halves; ++hh) {
      const int s = merged ? hh : l;
      TileEpi& e = j.bwd[hh];
      e.addend = static_cast<const __nv_bfloat16*>(d->addend[s]);
      e.addend2 = static_cast<const __nv_bfloat16*>(d->addend2[s]);
      e.mask_elu = static_cast<const __nv_bfloat16*>(d->mask_elu[s]);
      e.out_raw = static_cast<__nv_bfloat16*>(d->dx[s]);
      e.out_h = d->hin;
      e.out_w = d->win;
      e.out_c = C;
      e.pix_stride = 1;
      e.n_valid = C;
      e.flags = d->flags & VUB_CONV_RELU;
    }
    *rc = rows_job_launch(j, stream);
    if (*rc) return true;
  }
  return true;
}

}  // namespace vub
