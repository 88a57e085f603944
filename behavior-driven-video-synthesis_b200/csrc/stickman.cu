// GPU stickman rasteriser — the rendering front-end of frame synthesis.
//
// The reference renders every pose of a behavior-net sequence on the host: `make_joint_img` (lib/utils.py:325-513)
// draws the body polygon with cv2.fillPoly and the limbs with cv2.line into three uint8 planes, one frame at a
// time (data/data_conversions_3d.py:1154-1160), and `stickman_transforms` turns the HWC image into the network input.
// Here one launch rasterises a whole batch of poses into the uint8 HWC images that `vub_stem_fwd_u8` consumes
// directly.  Integer work, so the bar is bit-exactness: the two OpenCV primitives are restated exactly (thickness 1,
// 8-connected, shift 0; see oracle/stickman_oracle.py for the algorithm and how it is pinned to the reference):
//   line     : clipLine (Cohen-Sutherland, double-precision intersection, truncation) + left-to-right Bresenham with
//              err = dx - 2*dy;
//   fillPoly : outline lines + scan-line fill of 16.16 fixed-point edges built from the clipped end points (start
//              extrapolated back to the unclipped row; a degenerate clip leaves a vertical edge at the border),
//              spans [ceil(x_a), floor(x_b)] between consecutive sorted crossings.
// One block per frame: (0) clear, (1) thread 0 builds the frame's primitive list exactly in the reference's drawing
// order, (2) polygon outline + rows in parallel, (3) one thread per colour plane draws that plane's lines in order
// (later lines overwrite earlier ones like in the reference).
#include "common.h"

namespace vub {

constexpr int kStickThreads = 256;
constexpr int kMaxPrims = 96;
constexpr int kMaxPoly = 8;
constexpr long long XY_ONE = 1LL << 16;

struct LinePrim {
  int x0, y0, x1, y1, chan, color;
};

struct StickParams {
  const double* joints;  // [batch][n_joints][2] (x, y)
  vub_stickman_model m;
  int n_joints, h, w;
  uint8_t* out;          // [batch][h][w][3]
};

__device__ __forceinline__ bool clip_line_dev(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = w - 1, bottom = h - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += static_cast<long long>(static_cast<double>(a - y1) * static_cast<double>(x2 - x1) / static_cast<double>(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += static_cast<long long>(static_cast<double>(a - y2) * static_cast<double>(x2 - x1) / static_cast<double>(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += static_cast<long long>(static_cast<double>(a - x1) * static_cast<double>(y2 - y1) / static_cast<double>(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += static_cast<long long>(static_cast<double>(a - x2) * static_cast<double>(y2 - y1) / static_cast<double>(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// cv2.line(thickness 1) into colour plane `chan` (mask bit c set = also plane c) of one HWC frame
__device__ void draw_line_dev(uint8_t* img, int w, int h, long long ax, long long ay, long long bx, long long by, int chan_mask,
                              int color) {
  if (!clip_line_dev(w, h, ax, ay, bx, by)) return;
  long long dx = bx - ax, dy = by - ay;
  if (dx < 0) {
    long long t = ax; ax = bx; bx = t;
    t = ay; ay = by; by = t;
    dx = -dx;
    dy = -dy;
  }
  const int sy = dy >= 0 ? 1 : -1;
  if (dy < 0) dy = -dy;
  const bool steep = dy > dx;
  const long long major = steep ? dy : dx, minor = steep ? dx : dy;
  long long err = major - 2 * minor;
  const long long plus = 2 * major, minus = -2 * minor;
  int x = static_cast<int>(ax), y = static_cast<int>(ay);
  for (long long i = 0; i <= major; ++i) {
    uint8_t* p = img + (static_cast<size_t>(y) * w + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if ((chan_mask >> c) & 1) p[c] = static_cast<uint8_t>(color >> (8 * c));
    const bool m = err < 0;
    err += minus + (m ? plus : 0);
    if (steep) {
      y += sy;
      x += m ? 1 : 0;
    } else {
      x += 1;
      y += m ? sy : 0;
    }
  }
}

__device__ __forceinline__ long long trunc_div(long long num, long long den) { return num / den; }  // C++ truncates

__global__ void __launch_bounds__(kStickThreads) stickman_kernel(const StickParams p) {
  __shared__ LinePrim s_lines[kMaxPrims];
  __shared__ int s_n_lines;
  __shared__ int s_poly[kMaxPoly][2];
  __shared__ int s_n_poly;
  __shared__ long long s_edge[kMaxPoly][4];  // y0, y1, x (16.16 at row y0), dx
  __shared__ int s_n_edges;
  const int b = blockIdx.x;
  uint8_t* img = p.out + static_cast<size_t>(b) * p.h * p.w * 3;
  const double* J = p.joints + static_cast<size_t>(b) * p.n_joints * 2;
  const vub_stickman_model& m = p.m;

  // (0) clear the frame (16-byte stores; h * w * 3 is a multiple of 16 for every supported size)
  {
    uint4* q = reinterpret_cast<uint4*>(img);
    const size_t n16 = static_cast<size_t>(p.h) * p.w * 3 / 16;
    for (size_t i = threadIdx.x; i < n16; i += kStickThreads) q[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  // (1) primitive list, in the reference's drawing order (lib/utils.py:344-506)
  if (threadIdx.x == 0) {
    int nl = 0;
    auto jx = [&](int j) { return J[2 * j]; };
    auto jy = [&](int j) { return J[2 * j + 1]; };
    auto ok = [&](int j) { return jx(j) >= 0.0 && jy(j) >= 0.0; };
    auto add = [&](double ax, double ay, double bx, double by, int chan_mask, int color) {
      if (nl < kMaxPrims)
        s_lines[nl++] = LinePrim{static_cast<int>(static_cast<long long>(ax)), static_cast<int>(static_cast<long long>(ay)),
                                 static_cast<int>(static_cast<long long>(bx)), static_cast<int>(static_cast<long long>(by)),
                                 chan_mask, color};
    };
    int np_ = 0;
    if (m.n_body > 2) {
      for (int i = 0; i < m.n_body; ++i)
        if (ok(m.body[i])) {
          s_poly[np_][0] = static_cast<int>(static_cast<long long>(jx(m.body[i])));
          s_poly[np_][1] = static_cast<int>(static_cast<long long>(jy(m.body[i])));
          ++np_;
        }
      if (np_ <= 2) np_ = 0;
    }
    s_n_poly = np_;
    for (int i = 0; i < m.n_right; ++i) {
      const int a = m.right[i][0], c = m.right[i][1];
      if (ok(a) && ok(c)) add(jx(a), jy(a), jx(c), jy(c), m.colored ? 1 << m.right_chan[i] : 2, m.colored ? m.right_col[i] * 0x010101 : 0xffffff);
    }
    for (int i = 0; i < m.n_left; ++i) {
      const int a = m.left[i][0], c = m.left[i][1];
      if (ok(a) && ok(c)) add(jx(a), jy(a), jx(c), jy(c), m.colored ? 1 << m.left_chan[i] : 1, m.colored ? m.left_col[i] * 0x010101 : 0xffffff);
    }
    double throat_len = 0.0;
    if (m.n_head == 0) {
      const int rs = m.rshoulder, ls = m.lshoulder, cn = m.headup;
      double nx = -1.0, ny = -1.0;
      if (!(jx(rs) < 0.0 || jx(ls) < 0.0 || jy(rs) < 0.0 || jy(ls) < 0.0)) {
        nx = 0.5 * (jx(rs) + jx(ls));
        ny = 0.5 * (jy(rs) + jy(ls));
      }
      if (nx >= 0.0 && ny >= 0.0 && ok(cn)) {
        throat_len = sqrt((nx - jx(cn)) * (nx - jx(cn)) + (ny - jy(cn)) * (ny - jy(cn)));
        add(nx, ny, jx(cn), jy(cn), 3, 0x7f7f7f);  // planes 0 and 1, colour 127
      }
    } else {
      for (int i = 0; i < m.n_head; ++i) {
        const int a = m.head[i][0], c = m.head[i][1];
        if (ok(a) && ok(c)) {
          const double l = sqrt((jx(a) - jx(c)) * (jx(a) - jx(c)) + (jy(a) - jy(c)) * (jy(a) - jy(c)));
          throat_len = l > throat_len ? l : throat_len;
          add(jx(a), jy(a), jx(c), jy(c), m.colored ? 1 << m.head_chan[i] : 3, m.colored ? m.head_col[i] * 0x010101 : 0x7f7f7f);
        }
      }
    }
    for (int i = 0; i < m.n_face; ++i) {
      const int a = m.face[i][0], c = m.face[i][1];
      if (ok(a) && ok(c)) {
        const double l = sqrt((jx(a) - jx(c)) * (jx(a) - jx(c)) + (jy(a) - jy(c)) * (jy(a) - jy(c)));
        if (l < throat_len)  // the reference indexes line_colors[2] with the FACE loop's own counter
          add(jx(a), jy(a), jx(c), jy(c), m.colored ? 1 << m.head_chan[i] : 3, m.colored ? m.head_col[i] * 0x010101 : 0x7f7f7f);
      }
    }
    s_n_lines = nl;
    // polygon edges (CollectPolyEdges): slope from the clipped end points, start extrapolated to the unclipped row
    int ne = 0;
    for (int i = 0; i < np_; ++i) {
      const int i0 = i == 0 ? np_ - 1 : i - 1;
      const long long x0 = s_poly[i0][0], y0 = s_poly[i0][1], x1 = s_poly[i][0], y1 = s_poly[i][1];
      long long c0x = x0 << 16, c0y = y0, c1x = x1 << 16, c1y = y1;
      if (!(x0 >= 0 && x0 < p.w && x1 >= 0 && x1 < p.w && y0 >= 0 && y0 < p.h && y1 >= 0 && y1 < p.h)) {
        long long tx0 = x0, ty0 = y0, tx1 = x1, ty1 = y1;
        clip_line_dev(p.w, p.h, tx0, ty0, tx1, ty1);
        c0x = tx0 << 16;  // the x's are always the clipped ones; the y's only when the visible part spans > 1 row
        c1x = tx1 << 16;  // (a degenerate clip leaves a vertical edge at the image border over the original rows)
        if (ty0 != ty1) {
          c0y = ty0;
          c1y = ty1;
        }
      }
      if (y0 != y1) {
        const long long dxe = trunc_div(c1x - c0x, c1y - c0y);
        if (y0 < y1) {
          s_edge[ne][0] = y0; s_edge[ne][1] = y1; s_edge[ne][2] = c0x + (y0 - c0y) * dxe;
        } else {
          s_edge[ne][0] = y1; s_edge[ne][1] = y0; s_edge[ne][2] = c1x + (y1 - c1y) * dxe;
        }
        s_edge[ne][3] = dxe;
        ++ne;
      }
    }
    s_n_edges = ne;
  }
  __syncthreads();  // also orders the clear before every drawing store of this block
  // (2) body polygon: colour (0, 127, 255) -> plane 0 receives zeros (already clear), planes 1 and 2 are painted
  const int np_ = s_n_poly;
  if (np_ > 0) {
    if (threadIdx.x < np_) {  // outline, one edge per thread (same colour everywhere: order does not matter)
      const int i = threadIdx.x, i0 = i == 0 ? np_ - 1 : i - 1;
      draw_line_dev(img, p.w, p.h, s_poly[i0][0], s_poly[i0][1], s_poly[i][0], s_poly[i][1], 6, 0xff7f00);
    }
    const int ne = s_n_edges;
    long long y_min = 1LL << 40, y_max = -(1LL << 40);
    for (int e = 0; e < ne; ++e) {
      y_min = s_edge[e][0] < y_min ? s_edge[e][0] : y_min;
      y_max = s_edge[e][1] > y_max ? s_edge[e][1] : y_max;
    }
    if (y_max > p.h) y_max = p.h;
    if (y_min < 0) y_min = 0;
    for (long long y = y_min + threadIdx.x; y < y_max; y += kStickThreads) {
      long long xs[kMaxPoly];
      int n = 0;
      for (int e = 0; e < ne; ++e)
        if (s_edge[e][0] <= y && y < s_edge[e][1]) {
          long long x = s_edge[e][2] + (y - s_edge[e][0]) * s_edge[e][3];
          int k = n++;
          while (k > 0 && xs[k - 1] > x) {  // insertion sort
            xs[k] = xs[k - 1];
            --k;
          }
          xs[k] = x;
        }
      for (int k = 0; k + 1 < n; k += 2) {
        long long xa = (xs[k] + XY_ONE - 1) >> 16, xb = xs[k + 1] >> 16;
        if (xa < p.w && xb >= 0) {
          if (xa < 0) xa = 0;
          if (xb >= p.w) xb = p.w - 1;
          uint8_t* row = img + (static_cast<size_t>(y) * p.w) * 3;
          for (long long x = xa; x <= xb; ++x) {
            row[3 * x + 1] = 127;
            row[3 * x + 2] = 255;
          }
        }
      }
    }
  }
  __syncthreads();
  // (3) limbs: plane c is drawn by thread c, in list order (a later line overwrites an earlier one of the same plane)
  if (threadIdx.x < 3) {
    const int c = threadIdx.x, nl = s_n_lines;
    for (int i = 0; i < nl; ++i) {
      const LinePrim l = s_lines[i];
      if ((l.chan >> c) & 1) draw_line_dev(img, p.w, p.h, l.x0, l.y0, l.x1, l.y1, 1 << c, l.color);
    }
  }
}

}  // namespace vub

using namespace vub;

extern "C" int vub_stickman_render(const double* joints, int32_t batch, int32_t n_joints, const vub_stickman_model* model,
                                   int32_t height, int32_t width, uint8_t* out_hwc, void* stream) {
  clear_error();
  if (!joints || !model || !out_hwc) return fail(VUB_E_NULL, "stickman_render: NULL pointer");
  if (batch <= 0 || n_joints <= 0 || height <= 0 || width <= 0 || (static_cast<long long>(height) * width * 3) % 16)
    return fail(VUB_E_SHAPE, "stickman_render: batch %d joints %d size %dx%d (h*w*3 must be a multiple of 16)", batch,
                n_joints, height, width);
  if (!aligned16(out_hwc)) return fail(VUB_E_ALIGN, "stickman_render: output not 16B aligned");
  const vub_stickman_model& m = *model;
  if (m.n_body < 0 || m.n_body > kMaxPoly || m.n_right < 0 || m.n_right > 16 || m.n_left < 0 || m.n_left > 16 || m.n_head < 0 ||
      m.n_head > 8 || m.n_face < 0 || m.n_face > 16)
    return fail(VUB_E_SHAPE, "stickman_render: joint model too large (body <= 8, right/left/face <= 16, head <= 8 lines)");
  auto in_range = [&](int j) { return j >= 0 && j < n_joints; };
  bool ok = in_range(m.rshoulder) && in_range(m.lshoulder) && in_range(m.headup);
  for (int i = 0; i < m.n_body; ++i) ok = ok && in_range(m.body[i]);
  for (int i = 0; i < m.n_right; ++i) ok = ok && in_range(m.right[i][0]) && in_range(m.right[i][1]);
  for (int i = 0; i < m.n_left; ++i) ok = ok && in_range(m.left[i][0]) && in_range(m.left[i][1]);
  for (int i = 0; i < m.n_head; ++i) ok = ok && in_range(m.head[i][0]) && in_range(m.head[i][1]);
  for (int i = 0; i < m.n_face; ++i) ok = ok && in_range(m.face[i][0]) && in_range(m.face[i][1]);
  if (!ok) return fail(VUB_E_SHAPE, "stickman_render: joint index outside [0, %d)", n_joints);
  StickParams p;
  p.joints = joints;
  p.m = m;
  p.n_joints = n_joints;
  p.h = height;
  p.w = width;
  p.out = out_hwc;
  stickman_kernel<<<batch, kStickThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("stickman_render");
}
