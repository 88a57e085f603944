// Native GPU tests of the training-half kernels of libvunet_b200.so (vub_conv_dgrad, vub_conv_wgrad,
// vub_wn_bwd, vub_stem_bwd, vub_latent_bwd, the padded layout converter).  Every case compares the CUDA
// path with a plain CPU computation on the same bf16-rounded inputs.  Exit code = failed cases.
//   ./test_bwd        -> correctness        ./test_bwd perf -> plus timings of the dominant shapes
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vunet_b200.h"

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(100);                                                                    \
    }                                                                               \
  } while (0)

static uint32_t g_seed = 777u;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFFFF) / 8388608.0f - 1.0f;
}
static float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  explicit DevBuf(size_t n_) : n(n_ ? n_ : 8) { CK(cudaMalloc(&p, n * sizeof(T))); CK(cudaMemset(p, 0xFF, n * sizeof(T))); }
  ~DevBuf() { cudaFree(p); }
  void up(const std::vector<T>& h) { CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
  void zero() { CK(cudaMemset(p, 0, n * sizeof(T))); }
  std::vector<T> down() const {
    std::vector<T> h(n);
    CK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }
};
static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
  std::vector<__nv_bfloat16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
  return o;
}
static std::vector<float> rand_bf16(size_t n, float amp) {
  std::vector<float> v(n);
  for (auto& t : v) t = bf16r(amp * frand());
  return v;
}

struct Cmp {
  double num = 0, den = 0, maxabs = 0;
  size_t bad = 0, n = 0;
  void add(float got, float ref, float atol, float rtol) {
    const double d = (double)got - ref;
    num += d * d;
    den += (double)ref * ref;
    if (fabs(d) > maxabs) maxabs = fabs(d);
    if (!(fabs(d) <= atol + rtol * fabs(ref))) ++bad;
    ++n;
  }
  double rel() const { return sqrt(num / (den > 0 ? den : 1)); }
};
static int g_fail = 0;
static void report(const std::string& name, const Cmp& c) {
  const bool ok = c.bad == 0 && c.n > 0;
  printf("[%s] %-64s n=%zu relL2=%.3e maxabs=%.3e bad=%zu\n", ok ? "PASS" : "FAIL", name.c_str(), c.n, c.rel(),
         c.maxabs, c.bad);
  if (!ok) ++g_fail;
}
static void sync_or_die(const char* what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[FAIL] %s: sync error %s\n", what, cudaGetErrorString(e));
    exit(101);
  }
}

// ------------------------------------------------------------------------------------------------

// Tile-major, pre-swizzled copy of a K-major bf16 matrix [rows][K] (K % 64 == 0): block k/64, row, 16-byte chunk
// ((k%64)/8) ^ (row%8) — the shared-memory image of a SWIZZLE_128B operand tile (vub_wn_entry.w_tiled / wd_tiled).
static std::vector<float> tile_major(const std::vector<float>& w, int rows, int K) {
  std::vector<float> t(w.size());
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      const int kb = k >> 6, kc = k & 63;
      t[((size_t)kb * rows + r) * 64 + ((((kc >> 3) ^ (r & 7)) << 3) | (kc & 7))] = w[(size_t)r * K + k];
    }
  return t;
}
static int g_tiled = 0;  // 1: also hand the kernels the tile-major weights (1-D bulk weight loads)

struct BwdCase {
  const char* name;
  int B, H, W;  // conv INPUT extents
  int c_src, nsrc, cout_k, ksize, stride;
  bool d2s, mask, addend, addend2, inplace;
};

// gradient w.r.t. the conv output at conv-grid pixel (n, oh, ow), channel co
struct DyView {
  const std::vector<float>* dy;
  int Ho, Wo, cout_k;
  bool d2s;
  float at(int n, int oh, int ow, int co) const {
    if (!d2s) return (*dy)[(((size_t)n * Ho + oh) * Wo + ow) * cout_k + co];
    const int cq = cout_k / 4, q = co / cq, cp = co % cq;
    return (*dy)[(((size_t)n * 2 * Ho + 2 * oh + (q >> 1)) * 2 * Wo + 2 * ow + (q & 1)) * cq + cp];
  }
};

// VUB_SANITIZE=1 (tools/sanitize.sh): keep the cases below 16 k pixels under compute-sanitizer
static bool sanitize_skip(long long pixels) {
  static const bool on = getenv("VUB_SANITIZE") != nullptr;
  return on && pixels > 16384;
}

static void run_dgrad(const BwdCase& c) {
  if (sanitize_skip((long long)c.B * c.H * c.W)) return;
  const int taps = c.ksize * c.ksize, pad = c.ksize / 2;
  const int Ho = c.H / c.stride, Wo = c.W / c.stride;
  const size_t dy_elems = (size_t)c.B * Ho * Wo * c.cout_k;
  const size_t x_elems = (size_t)c.B * c.H * c.W * c.c_src;
  const int N = c.nsrc * c.c_src, K = taps * c.cout_k;
  std::vector<float> dy = rand_bf16(dy_elems, 1.0f), wd = rand_bf16((size_t)N * K, 0.06f);
  std::vector<float> mask[2], add1[2], add2[2];
  for (int s = 0; s < c.nsrc; ++s) {
    if (c.mask) mask[s] = rand_bf16(x_elems, 1.0f);  // plays ELU(x): negative values give ELU' = e + 1
    if (c.addend) add1[s] = rand_bf16(x_elems, 1.0f);
    if (c.addend2) add2[s] = rand_bf16(x_elems, 1.0f);
  }
  DyView dv{&dy, Ho, Wo, c.cout_k, c.d2s};
  std::vector<float> ref[2];
  for (int s = 0; s < c.nsrc; ++s) {
    ref[s].assign(x_elems, 0.f);
    for (int n = 0; n < c.B; ++n)
      for (int ih = 0; ih < c.H; ++ih)
        for (int iw = 0; iw < c.W; ++iw) {
          std::vector<double> acc(c.c_src, 0.0);
          for (int t = 0; t < taps; ++t) {
            const int r = t / c.ksize, q = t % c.ksize;
            const int nh = ih + pad - r, nw = iw + pad - q;  // = stride * oh
            if (nh % c.stride || nw % c.stride) continue;
            const int oh = nh / c.stride, ow = nw / c.stride;
            if (nh < 0 || nw < 0 || oh >= Ho || ow >= Wo) continue;
            for (int co = 0; co < c.cout_k; ++co) {
              const float g = dv.at(n, oh, ow, co);
              const float* wp = &wd[(size_t)(s * c.c_src) * K + (size_t)t * c.cout_k + co];
              for (int ci = 0; ci < c.c_src; ++ci) acc[ci] += (double)g * wp[(size_t)ci * K];
            }
          }
          const size_t o = (((size_t)n * c.H + ih) * c.W + iw) * c.c_src;
          for (int ci = 0; ci < c.c_src; ++ci) {
            float y = (float)acc[ci];
            if (c.mask) {
              const float e = mask[s][o + ci];
              y *= e > 0.f ? 1.0f : e + 1.0f;
            }
            if (c.addend) y += add1[s][o + ci];
            if (c.addend2) y += add2[s][o + ci];
            ref[s][o + ci] = y;
          }
        }
  }
  DevBuf<__nv_bfloat16> d_dy(dy_elems), d_wd((size_t)N * K);
  d_dy.up(to_bf16(dy));
  d_wd.up(to_bf16(wd));
  DevBuf<__nv_bfloat16> d_wdt((size_t)N * K);
  const bool tiled = g_tiled && c.cout_k % 64 == 0;
  if (tiled) d_wdt.up(to_bf16(tile_major(wd, N, K)));
  DevBuf<__nv_bfloat16>*d_dx[2] = {nullptr, nullptr}, *d_m[2] = {nullptr, nullptr}, *d_a1[2] = {nullptr, nullptr},
                       *d_a2[2] = {nullptr, nullptr};
  vub_dgrad_desc d;
  memset(&d, 0, sizeof(d));
  for (int s = 0; s < c.nsrc; ++s) {
    d_dx[s] = new DevBuf<__nv_bfloat16>(x_elems);
    if (c.mask) { d_m[s] = new DevBuf<__nv_bfloat16>(x_elems); d_m[s]->up(to_bf16(mask[s])); d.mask_elu[s] = d_m[s]->p; }
    if (c.addend) {
      if (c.inplace) { d_dx[s]->up(to_bf16(add1[s])); d.addend[s] = d_dx[s]->p; }
      else { d_a1[s] = new DevBuf<__nv_bfloat16>(x_elems); d_a1[s]->up(to_bf16(add1[s])); d.addend[s] = d_a1[s]->p; }
    }
    if (c.addend2) { d_a2[s] = new DevBuf<__nv_bfloat16>(x_elems); d_a2[s]->up(to_bf16(add2[s])); d.addend2[s] = d_a2[s]->p; }
    d.dx[s] = d_dx[s]->p;
  }
  d.dy = d_dy.p;
  d.wd = d_wd.p;
  d.wd_tiled = tiled ? d_wdt.p : nullptr;
  d.nsrc = c.nsrc; d.c_src = c.c_src; d.batch = c.B; d.hin = c.H; d.win = c.W; d.cout_k = c.cout_k;
  d.ksize = c.ksize; d.stride = c.stride; d.flags = c.d2s ? VUB_CONV_D2S : 0;
  const int rc = vub_conv_dgrad(&d, nullptr);
  if (rc) {
    printf("[FAIL] dgrad %-58s launch rc=%d (%s)\n", c.name, rc, vub_last_error());
    ++g_fail;
  } else {
    sync_or_die(c.name);
    for (int s = 0; s < c.nsrc; ++s) {
      auto got = d_dx[s]->down();
      Cmp cm;
      for (size_t i = 0; i < x_elems; ++i) cm.add(__bfloat162float(got[i]), ref[s][i], 2e-2f, 1e-2f);
      report(std::string("dgrad ") + c.name + (s ? " src1" : " src0"), cm);
    }
  }
  for (int s = 0; s < 2; ++s) { delete d_dx[s]; delete d_m[s]; delete d_a1[s]; delete d_a2[s]; }
}

static void run_wgrad(const BwdCase& c) {
  if (sanitize_skip((long long)c.B * c.H * c.W)) return;
  const int taps = c.ksize * c.ksize, pad = c.ksize / 2;
  const int Ho = c.H / c.stride, Wo = c.W / c.stride;
  const size_t dy_elems = (size_t)c.B * Ho * Wo * c.cout_k;
  const size_t x_elems = (size_t)c.B * c.H * c.W * c.c_src;
  const int KR = c.nsrc * taps * c.c_src;
  const int pitch = KR + 32;  // exercise gt_pitch > K
  std::vector<float> dy = rand_bf16(dy_elems, 1.0f), x[2];
  for (int s = 0; s < c.nsrc; ++s) x[s] = rand_bf16(x_elems, 1.0f);
  DyView dv{&dy, Ho, Wo, c.cout_k, c.d2s};
  std::vector<double> ref((size_t)KR * c.cout_k, 0.0), ref_sum(c.cout_k, 0.0);
  for (int n = 0; n < c.B; ++n)
    for (int oh = 0; oh < Ho; ++oh)
      for (int ow = 0; ow < Wo; ++ow) {
        std::vector<float> g(c.cout_k);
        for (int co = 0; co < c.cout_k; ++co) {
          g[co] = dv.at(n, oh, ow, co);
          ref_sum[co] += g[co];
        }
        for (int s = 0; s < c.nsrc; ++s)
          for (int t = 0; t < taps; ++t) {
            const int r = t / c.ksize, q = t % c.ksize;
            const int ih = oh * c.stride + r - pad, iw = ow * c.stride + q - pad;
            if (ih < 0 || ih >= c.H || iw < 0 || iw >= c.W) continue;
            const float* xp = &x[s][(((size_t)n * c.H + ih) * c.W + iw) * c.c_src];
            for (int ci = 0; ci < c.c_src; ++ci) {
              double* rp = &ref[(size_t)((s * taps + t) * c.c_src + ci) * c.cout_k];
              const double xv = xp[ci];
              for (int co = 0; co < c.cout_k; ++co) rp[co] += xv * g[co];
            }
          }
      }
  DevBuf<__nv_bfloat16> d_dy(dy_elems), d_x0(x_elems), d_x1(c.nsrc == 2 ? x_elems : 8);
  d_dy.up(to_bf16(dy));
  d_x0.up(to_bf16(x[0]));
  if (c.nsrc == 2) d_x1.up(to_bf16(x[1]));
  vub_wgrad_desc d;
  memset(&d, 0, sizeof(d));
  d.nsrc = c.nsrc; d.c_src = c.c_src; d.batch = c.B; d.hin = c.H; d.win = c.W; d.cout_k = c.cout_k; d.gt_pitch = pitch;
  d.ksize = c.ksize; d.stride = c.stride; d.flags = c.d2s ? VUB_CONV_D2S : 0;
  const int ns = vub_conv_wgrad_splits(&d);
  if (ns <= 0) {
    printf("[FAIL] wgrad %-58s splits rc=%d (%s)\n", c.name, ns, vub_last_error());
    ++g_fail;
    return;
  }
  const size_t slice = (size_t)c.cout_k * pitch + 64;  // exercise split_stride > cout_k * pitch
  DevBuf<float> d_gt(slice * ns), d_sum((size_t)(c.cout_k + 8) * ns);
  CK(cudaMemset(d_gt.p, 0, d_gt.n * sizeof(float)));  // pad regions must stay untouched (checked below)
  d.dy = d_dy.p; d.x[0] = d_x0.p; d.x[1] = c.nsrc == 2 ? d_x1.p : nullptr; d.gt = d_gt.p; d.dy_sum = d_sum.p;
  d.n_split = ns; d.split_stride = (int64_t)slice; d.sum_stride = c.cout_k + 8;
  const int rc = vub_conv_wgrad(&d, nullptr);
  if (rc) {
    printf("[FAIL] wgrad %-58s launch rc=%d (%s)\n", c.name, rc, vub_last_error());
    ++g_fail;
    return;
  }
  sync_or_die(c.name);
  auto gt = d_gt.down();
  auto sm = d_sum.down();
  const float scale = sqrtf((float)c.B * Ho * Wo);  // typical magnitude of a correlation of +-1 noise
  Cmp cg, cs, cz;
  for (int co = 0; co < c.cout_k; ++co) {
    for (int k = 0; k < KR; ++k) {
      double g = 0;
      for (int sp = 0; sp < ns; ++sp) g += gt[sp * slice + (size_t)co * pitch + k];
      cg.add((float)g, (float)ref[(size_t)k * c.cout_k + co], 2e-3f * scale, 2e-3f);
    }
    for (int sp = 0; sp < ns; ++sp)
      for (int k = KR; k < pitch; ++k) cz.add(gt[sp * slice + (size_t)co * pitch + k], 0.f, 0.f, 0.f);
  }
  for (int co = 0; co < c.cout_k; ++co) {
    double g = 0;
    for (int sp = 0; sp < ns; ++sp) g += sm[(size_t)sp * (c.cout_k + 8) + co];
    cs.add((float)g, (float)ref_sum[co], 2e-3f * scale, 2e-3f);
  }
  report(std::string("wgrad ") + c.name + " gt", cg);
  report(std::string("wgrad ") + c.name + " dy_sum", cs);
  report(std::string("wgrad ") + c.name + " pad columns untouched", cz);
}

// ------------------------------------------------------------------------------------------------
static void test_wn_prep_wd() {
  // forward packing is covered by test_kernels; here: the data-gradient matrix wd of a merged pair
  const int cout = 64, cin = 128, k = 3, c_src = 64, taps = 9, len = cin * taps, wd_cout = 128;
  std::vector<float> v[2], g[2], ga[2], be[2], bi[2];
  DevBuf<__nv_bfloat16> d_w((size_t)2 * cout * len), d_wd((size_t)cin * taps * wd_cout);
  d_wd.zero();
  DevBuf<float> d_sc(2 * cout), d_sh(2 * cout);
  std::vector<DevBuf<float>*> bufs;
  vub_wn_entry tab[2];
  for (int l = 0; l < 2; ++l) {
    v[l].resize((size_t)cout * len);
    for (auto& t : v[l]) t = frand();
    g[l].resize(cout); ga[l].resize(cout); be[l].resize(cout); bi[l].resize(cout);
    for (int i = 0; i < cout; ++i) { g[l][i] = 0.5f + fabsf(frand()); ga[l][i] = 0.7f + 0.6f * fabsf(frand()); be[l][i] = 0.1f * frand(); bi[l][i] = 0.1f * frand(); }
    DevBuf<float>* b[5] = {new DevBuf<float>(v[l].size()), new DevBuf<float>(cout), new DevBuf<float>(cout), new DevBuf<float>(cout), new DevBuf<float>(cout)};
    b[0]->up(v[l]); b[1]->up(g[l]); b[2]->up(ga[l]); b[3]->up(be[l]); b[4]->up(bi[l]);
    for (auto* q : b) bufs.push_back(q);
    vub_wn_entry& e = tab[l];
    memset(&e, 0, sizeof(e));
    e.v = b[0]->p; e.g = b[1]->p; e.gamma = b[2]->p; e.beta = b[3]->p; e.bias = b[4]->p;
    e.w_packed = d_w.p + (size_t)l * cout * len; e.scale = d_sc.p + l * cout; e.shift = d_sh.p + l * cout;
    e.cout = cout; e.cin = cin; e.ktaps = taps; e.c_src = c_src; e.ktot = len; e.row_start = l * cout;
    e.wd = d_wd.p; e.wd_cout = wd_cout; e.wd_col0 = l * cout;
  }
  vub_wn_entry* dtab;
  CK(cudaMalloc(&dtab, sizeof(tab)));
  CK(cudaMemcpy(dtab, tab, sizeof(tab), cudaMemcpyHostToDevice));
  const int rc = vub_wn_prep(dtab, 2, 2 * cout, nullptr);
  if (rc) { printf("[FAIL] wn_prep(wd) rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  sync_or_die("wn_prep wd");
  auto wd = d_wd.down();
  Cmp c;
  for (int l = 0; l < 2; ++l)
    for (int co = 0; co < cout; ++co) {
      double ss = 0;
      for (int i = 0; i < len; ++i) ss += (double)v[l][(size_t)co * len + i] * v[l][(size_t)co * len + i];
      const double sc = ga[l][co] * g[l][co] / sqrt(ss);
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < taps; ++t)
          c.add(__bfloat162float(wd[(size_t)ci * taps * wd_cout + t * wd_cout + l * cout + co]),
                (float)(v[l][((size_t)co * cin + ci) * taps + t] * sc), 1e-4f, 8e-3f);
    }
  report("wn_prep dgrad weights wd (merged pair, k order (tap, co))", c);
  cudaFree(dtab);
  for (auto* q : bufs) delete q;
  (void)k;
}

static void test_wn_bwd() {
  struct L { int cout, cin, k, c_src; };
  const L layers[] = {{32, 64, 3, 32}, {128, 256, 1, 128}, {3, 32, 3, 32}, {32, 30, 1, 30}};
  const int nl = 4;
  std::vector<vub_wnb_entry> tab(nl);
  std::vector<std::vector<float>> v(nl), g(nl), ga(nl), bi(nl), gt(nl), sm(nl);
  std::vector<std::vector<DevBuf<float>*>> bufs(nl);
  int rows = 0;
  for (int l = 0; l < nl; ++l) {
    const L& q = layers[l];
    const int taps = q.k * q.k, len = q.cin * taps, pitch = len + 4, col0 = 3;
    v[l].resize((size_t)q.cout * len);
    for (auto& t : v[l]) t = frand();
    g[l].resize(q.cout); ga[l].resize(q.cout); bi[l].resize(q.cout);
    for (int i = 0; i < q.cout; ++i) { g[l][i] = 0.5f + fabsf(frand()); ga[l][i] = 0.7f + 0.6f * fabsf(frand()); bi[l][i] = 0.1f * frand(); }
    const int ns = l + 1;  // 1..4 pixel splits, summed by the kernel
    const size_t sstride = (size_t)(q.cout + col0) * pitch + 12;
    gt[l].resize(sstride * ns);
    for (auto& t : gt[l]) t = 3.f * frand();
    sm[l].resize((size_t)(q.cout + col0 + 2) * ns);
    for (auto& t : sm[l]) t = 5.f * frand();
    auto mk = [&](const std::vector<float>& h) { auto* b = new DevBuf<float>(h.size()); b->up(h); bufs[l].push_back(b); return b->p; };
    auto out = [&](size_t n) { auto* b = new DevBuf<float>(n); bufs[l].push_back(b); return b->p; };
    vub_wnb_entry& e = tab[l];
    memset(&e, 0, sizeof(e));
    e.v = mk(v[l]); e.g = mk(g[l]); e.gamma = mk(ga[l]); e.bias = mk(bi[l]); e.gt = mk(gt[l]); e.dy_sum = mk(sm[l]);
    e.d_v = out(v[l].size()); e.d_g = out(q.cout); e.d_gamma = out(q.cout); e.d_beta = out(q.cout); e.d_bias = out(q.cout);
    e.cout = q.cout; e.cin = q.cin; e.ktaps = taps; e.c_src = q.c_src; e.gt_pitch = pitch; e.gt_col0 = col0; e.row_start = rows;
    e.n_split = ns; e.split_stride = (int64_t)sstride; e.sum_stride = q.cout + col0 + 2;
    rows += q.cout;
  }
  vub_wnb_entry* dtab;
  CK(cudaMalloc(&dtab, sizeof(vub_wnb_entry) * nl));
  CK(cudaMemcpy(dtab, tab.data(), sizeof(vub_wnb_entry) * nl, cudaMemcpyHostToDevice));
  const int rc = vub_wn_bwd(dtab, nl, rows, 0, 0, nullptr);
  if (rc) { printf("[FAIL] wn_bwd rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  sync_or_die("wn_bwd");
  Cmp cv, cs;
  for (int l = 0; l < nl; ++l) {
    const L& q = layers[l];
    const int taps = q.k * q.k, len = q.cin * taps, pitch = len + 4, col0 = 3;
    const int ns = l + 1;
    const size_t sstride = (size_t)(q.cout + col0) * pitch + 12;
    auto dv = bufs[l][6]->down(); auto dg = bufs[l][7]->down(); auto dga = bufs[l][8]->down();
    auto dbe = bufs[l][9]->down(); auto dbi = bufs[l][10]->down();
    for (int co = 0; co < q.cout; ++co) {
      double ss = 0, vg = 0;
      auto G = [&](int i) {
        const int ci = i / taps, t = i % taps, s = ci / q.c_src, cs_ = ci % q.c_src;
        double g = 0;
        for (int sp = 0; sp < ns; ++sp) g += gt[l][sp * sstride + (size_t)(col0 + co) * pitch + (s * taps + t) * q.c_src + cs_];
        return g;
      };
      for (int i = 0; i < len; ++i) { const double x = v[l][(size_t)co * len + i]; ss += x * x; vg += x * G(i); }
      double S = 0;
      for (int sp = 0; sp < ns; ++sp) S += sm[l][(size_t)sp * (q.cout + col0 + 2) + col0 + co];
      const double inv = 1.0 / sqrt(ss), dot = vg * inv;
      for (int i = 0; i < len; ++i)
        cv.add(dv[(size_t)co * len + i], (float)(ga[l][co] * g[l][co] * inv * (G(i) - v[l][(size_t)co * len + i] * inv * dot)), 1e-4f, 1e-4f);
      cs.add(dg[co], (float)(ga[l][co] * dot), 1e-4f, 1e-4f);
      cs.add(dga[co], (float)(g[l][co] * dot + bi[l][co] * S), 1e-4f, 1e-4f);
      cs.add(dbe[co], (float)S, 1e-5f, 1e-5f);
      cs.add(dbi[co], (float)(ga[l][co] * S), 1e-5f, 1e-5f);
    }
  }
  report("wn_bwd d_v (4 layers incl. split sources, stem, out_conv)", cv);
  report("wn_bwd d_g / d_gamma / d_beta / d_bias", cs);
  cudaFree(dtab);
  for (auto& bl : bufs) for (auto* b : bl) delete b;
}

static void test_stem_bwd(int B, int cx, int H, int W) {
  const int cout = 32;
  std::vector<float> x((size_t)B * cx * H * W);
  for (auto& t : x) t = frand();
  std::vector<float> dy = rand_bf16((size_t)B * H * W * cout, 1.0f);
  DevBuf<float> dx(x.size()), dgt((size_t)cx * cout), dsum(cout);
  DevBuf<__nv_bfloat16> ddy(dy.size());
  dx.up(x); ddy.up(to_bf16(dy)); dgt.zero(); dsum.zero();
  const int rc = vub_stem_bwd(dx.p, ddy.p, B, cx, H, W, cout, dgt.p, dsum.p, nullptr);
  if (rc) { printf("[FAIL] stem_bwd rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  sync_or_die("stem_bwd");
  auto gt = dgt.down(); auto sm = dsum.down();
  const float scale = sqrtf((float)B * H * W);
  Cmp c;
  for (int co = 0; co < cout; ++co) {
    double s = 0;
    for (size_t p = 0; p < (size_t)B * H * W; ++p) s += dy[p * cout + co];
    c.add(sm[co], (float)s, 1e-3f * scale, 1e-3f);
    for (int ci = 0; ci < cx; ++ci) {
      double a = 0;
      for (int n = 0; n < B; ++n)
        for (int p = 0; p < H * W; ++p) a += (double)x[((size_t)n * cx + ci) * H * W + p] * dy[((size_t)n * H * W + p) * cout + co];
      c.add(gt[co * cx + ci], (float)a, 1e-3f * scale, 1e-3f);
    }
  }
  char nm[96];
  snprintf(nm, sizeof nm, "stem_bwd B%d cx%d %dx%d", B, cx, H, W);
  report(nm, c);
}

static void test_latent_bwd(int B, int C, int h, int w, bool with_dz, bool with_user) {
  const size_t n = (size_t)B * C * h * w;
  std::vector<float> ls(n), eps(n), dmu(n), dls(n);
  for (auto& t : ls) t = 0.5f + 0.45f * frand();
  for (auto& t : eps) t = 1.5f * frand();
  for (auto& t : dmu) t = frand();
  for (auto& t : dls) t = frand();
  std::vector<float> dz = rand_bf16(n, 1.0f);  // NHWC
  DevBuf<float> d_ls(n), d_eps(n), d_dmu(n), d_dls(n);
  DevBuf<__nv_bfloat16> d_dz(n), d_out(2 * n);
  d_ls.up(ls); d_eps.up(eps); d_dmu.up(dmu); d_dls.up(dls); d_dz.up(to_bf16(dz));
  const int rc = vub_latent_bwd(nullptr, d_ls.p, d_eps.p, with_user ? d_dmu.p : nullptr, with_user ? d_dls.p : nullptr,
                                with_dz ? d_dz.p : nullptr, B, C, h, w, d_out.p, nullptr);
  if (rc) { printf("[FAIL] latent_bwd rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  sync_or_die("latent_bwd");
  auto out = d_out.down();
  Cmp c;
  for (int b = 0; b < B; ++b)
    for (int ch = 0; ch < C; ++ch)
      for (int p = 0; p < h * w; ++p) {
        const size_t o = ((size_t)b * C + ch) * h * w + p, oh = ((size_t)b * h * w + p);
        const float z = with_dz ? dz[oh * C + ch] : 0.f;
        const float gm = (with_user ? dmu[o] : 0.f) + z;
        const float gl = ((with_user ? dls[o] : 0.f) + z * eps[o] * expf(ls[o])) * ls[o] * (1.f - ls[o]);
        c.add(__bfloat162float(out[oh * 2 * C + ch]), gm, 4e-3f, 8e-3f);
        c.add(__bfloat162float(out[oh * 2 * C + C + ch]), gl, 4e-3f, 8e-3f);
      }
  char nm[96];
  snprintf(nm, sizeof nm, "latent_bwd B%d C%d %dx%d dz=%d user=%d", B, C, h, w, (int)with_dz, (int)with_user);
  report(nm, c);
}

static void test_pad_convert(int B, int C, int H, int W, int cpad) {
  std::vector<float> x((size_t)B * C * H * W);
  for (auto& t : x) t = 2.f * frand();
  DevBuf<float> dx(x.size());
  DevBuf<__nv_bfloat16> dd((size_t)B * H * W * cpad);
  dx.up(x);
  const int rc = vub_nchw_f32_to_nhwc_bf16_pad(dx.p, B, C, H, W, cpad, dd.p, nullptr);
  if (rc) { printf("[FAIL] pad convert rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  sync_or_die("pad convert");
  auto got = dd.down();
  Cmp c;
  for (int b = 0; b < B; ++b)
    for (int p = 0; p < H * W; ++p)
      for (int ch = 0; ch < cpad; ++ch)
        c.add(__bfloat162float(got[((size_t)b * H * W + p) * cpad + ch]), ch < C ? bf16r(x[((size_t)b * C + ch) * H * W + p]) : 0.f, 0.f, 0.f);
  report("padded NCHW fp32 -> NHWC bf16 converter", c);
}

// ------------------------------------------------------------------------------------------------
static void perf_bwd(const char* name, int B, int H, int W, int c_src, int nsrc, int cout_k, int ksize, int stride, bool d2s) {
  const int taps = ksize * ksize, Ho = H / stride, Wo = W / stride;
  const size_t dy_elems = (size_t)B * Ho * Wo * cout_k, x_elems = (size_t)B * H * W * c_src;
  const int KR = nsrc * taps * c_src;
  DevBuf<__nv_bfloat16> d_dy(dy_elems), d_x0(x_elems), d_x1(nsrc == 2 ? x_elems : 8), d_wd((size_t)nsrc * c_src * taps * cout_k);
  DevBuf<__nv_bfloat16> d_dx0(x_elems), d_dx1(nsrc == 2 ? x_elems : 8);
  vub_wgrad_desc wg;
  memset(&wg, 0, sizeof(wg));
  wg.nsrc = nsrc; wg.c_src = c_src; wg.batch = B; wg.hin = H; wg.win = W; wg.cout_k = cout_k; wg.gt_pitch = KR;
  wg.ksize = ksize; wg.stride = stride; wg.flags = d2s ? VUB_CONV_D2S : 0;
  const int ns = vub_conv_wgrad_splits(&wg) > 0 ? vub_conv_wgrad_splits(&wg) : 1;
  DevBuf<float> d_gt((size_t)KR * cout_k * ns), d_sum((size_t)cout_k * ns);
  d_dy.zero(); d_x0.zero(); d_x1.zero(); d_wd.zero();
  wg.dy = d_dy.p; wg.x[0] = d_x0.p; wg.x[1] = nsrc == 2 ? d_x1.p : nullptr; wg.gt = d_gt.p; wg.dy_sum = d_sum.p;
  wg.n_split = ns; wg.split_stride = (int64_t)KR * cout_k; wg.sum_stride = cout_k;
  vub_dgrad_desc dg;
  memset(&dg, 0, sizeof(dg));
  dg.dy = d_dy.p; dg.wd = d_wd.p; dg.dx[0] = d_dx0.p; dg.dx[1] = nsrc == 2 ? d_dx1.p : nullptr;
  dg.mask_elu[0] = d_x0.p; dg.mask_elu[1] = nsrc == 2 ? d_x1.p : nullptr;
  dg.nsrc = nsrc; dg.c_src = c_src; dg.batch = B; dg.hin = H; dg.win = W; dg.cout_k = cout_k; dg.ksize = ksize;
  dg.stride = stride; dg.flags = d2s ? VUB_CONV_D2S : 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double flop = 2.0 * B * Ho * Wo * (double)cout_k * KR;
  for (int which = 0; which < 3; ++which) {
    if (which == 2) {  // data gradient without the ELU' mask (epilogue reads nothing)
      dg.mask_elu[0] = nullptr;
      dg.mask_elu[1] = nullptr;
    }
    auto go = [&]() { return which == 0 ? vub_conv_wgrad(&wg, nullptr) : vub_conv_dgrad(&dg, nullptr); };
    int rc = 0;
    for (int i = 0; i < 3; ++i) rc |= go();
    if (rc) { printf("[PERF] %s %s: rc=%d %s\n", which ? "dgrad" : "wgrad", name, rc, vub_last_error()); continue; }
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) go();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double bytes = which == 0 ? 2.0 * (dy_elems + nsrc * (double)x_elems)
                                    : 2.0 * (dy_elems + (which == 1 ? 2.0 : 1.0) * nsrc * (double)x_elems);
    printf("[PERF] %s %-40s %8.3f ms  %7.1f TFLOP/s  %7.1f GB/s(algorithmic)\n",
           which == 0 ? "wgrad" : (which == 1 ? "dgrad" : "dgrad(no mask)"), name, ms,
           flop / ms * 1e-9, bytes / ms * 1e-6);
  }
}

int main(int argc, char** argv) {
  const bool perf = argc > 1 && !strcmp(argv[1], "perf");
  if (!perf && !getenv("VUB_ENABLE_SLAB")) setenv("VUB_ENABLE_SLAB", "1", 1);  // correctness runs cover the opt-in slab kernel
  if (argc >= 11 && !strcmp(argv[1], "layer")) {
    // layer B H W c_src nsrc cout_k k stride d2s  -> time wgrad + dgrad of one layer shape (for ncu captures)
    int a[9];
    for (int i = 0; i < 9; ++i) a[i] = atoi(argv[2 + i]);
    perf_bwd("layer", a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8] != 0);
    return 0;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s  SMs=%d  abi=%d\n", prop.name, prop.multiProcessorCount, vub_abi_version());
  //            name                                         B  H   W   c  ns cout k  s  d2s   mask  add   add2  inplace
  const BwdCase cases[] = {
      {"3x3 128->128 @16x16 B2 plain",                        2, 16, 16, 128, 1, 128, 3, 1, false, false, false, false, false},
      {"3x3 128->128 @16x16 B2 mask+addend+addend2 inplace",  2, 16, 16, 128, 1, 128, 3, 1, false, true, true, true, true},
      {"3x3 [128|128]->128 @8x8 B3 cat mask+addend2",         3, 8, 8, 128, 2, 128, 3, 1, false, true, false, true, false},
      {"3x3 [64|64]->64 @32x32 B2 cat mask+add+add2",         2, 32, 32, 64, 2, 64, 3, 1, false, true, true, true, true},
      {"3x3 [32|32]->32 @16x128 B1 cat mask+add2",            1, 16, 128, 32, 2, 32, 3, 1, false, true, false, true, false},
      {"3x3 32->32 @8x256 B1 mask+addend2",                   1, 8, 256, 32, 1, 32, 3, 1, false, true, false, true, false},
      {"1x1 128->128 @8x8 B2 mask+addend inplace",            2, 8, 8, 128, 1, 128, 1, 1, false, true, true, false, true},
      {"1x1 [128|128]->128 @4x4 B3 cat mask",                 3, 4, 4, 128, 2, 128, 1, 1, false, true, false, false, false},
      {"1x1 32->32 @16x64 B2 mask+addend",                    2, 16, 64, 32, 1, 32, 1, 1, false, true, true, false, false},
      {"3x3 128->256 (latent pair) @4x4 B3 addend",           3, 4, 4, 128, 1, 256, 3, 1, false, false, true, false, true},
      {"3x3 32->3(32) out_conv @16x64 B2",                    2, 16, 64, 32, 1, 32, 3, 1, false, false, false, false, false},
      {"3x3 s2 32->64 @32x64 B2 addend inplace",              2, 32, 64, 32, 1, 64, 3, 2, false, false, true, false, true},
      {"3x3 s2 64->128 @16x16 B2",                            2, 16, 16, 64, 1, 128, 3, 2, false, false, false, false, false},
      {"3x3 s2 128->128 @4x4 B3 addend",                      3, 4, 4, 128, 1, 128, 3, 2, false, false, true, false, false},
      {"3x3 128->512 D2S @8x8 B2",                            2, 8, 8, 128, 1, 512, 3, 1, true, false, false, false, false},
      {"3x3 128->256 D2S @16x16 B1 addend",                   1, 16, 16, 128, 1, 256, 3, 1, true, false, true, false, false},
      {"3x3 64->128 D2S @32x32 B1",                           1, 32, 32, 64, 1, 128, 3, 1, true, false, false, false, false},
  };
  // 256 / 512 channels per source (VunetOrg's 4-group prior, models/vunets.py:655-676): one launch per 128-channel slice
  const BwdCase wide_cases[] = {
      {"wide 3x3 512->128 (auto_lp) @4x4 B3 addend",          3, 4, 4, 512, 1, 128, 3, 1, false, false, true, false, false},
      {"wide 3x3 [512|512]->512 @4x4 B3 cat mask+add+add2",   3, 4, 4, 512, 2, 512, 3, 1, false, true, true, true, true},
      {"wide 3x3 [512|512]->512 @2x2 B9 cat mask+add2",       9, 2, 2, 512, 2, 512, 3, 1, false, true, false, true, false},
      {"wide 3x3 [256|256]->256 @1x1 B5 cat mask",            5, 1, 1, 256, 2, 256, 3, 1, false, true, false, false, false},
      {"wide 3x3 256->64 (auto_lp) @2x2 B3",                  3, 2, 2, 256, 1, 64, 3, 1, false, false, false, false, false},
      {"1x1 128->512 (nin of the prior blocks) @4x4 B3 mask", 3, 4, 4, 128, 1, 512, 1, 1, false, true, false, false, false},
      {"wide 3x3 512->128 @8x8 B5 mask+addend (5 k-steps)",   5, 8, 8, 512, 1, 128, 3, 1, false, true, true, false, false},
  };
  if (argc > 1 && !strcmp(argv[1], "wide")) {
    for (const BwdCase& c : wide_cases) run_dgrad(c);
    for (const BwdCase& c : wide_cases) run_wgrad(c);
    printf("%s: %d failed\n", g_fail ? "FAILED" : "ALL PASSED", g_fail);
    return g_fail;
  }
  for (const BwdCase& c : wide_cases) run_dgrad(c);
  for (const BwdCase& c : wide_cases) run_wgrad(c);
  for (const BwdCase& c : cases) run_dgrad(c);
  g_tiled = 1;  // the same data gradients with 1-D bulk weight loads from the tile-major wd
  for (const BwdCase& c : cases)
    if (c.cout_k % 64 == 0) run_dgrad(c);
  g_tiled = 0;
  for (const BwdCase& c : cases) run_wgrad(c);
  // shapes the activation-slab kernel takes over (data gradient with mirrored taps)
  const BwdCase slab_cases[] = {
      {"slab 3x3 128->128 @32x32 B2 mask+add+add2 inplace",   2, 32, 32, 128, 1, 128, 3, 1, false, true, true, true, true},
      {"slab 3x3 [128|128]->128 @64x64 B1 cat mask+add2",     1, 64, 64, 128, 2, 128, 3, 1, false, true, false, true, false},
      {"slab 3x3 128->256 (wide dy) @32x32 B2 addend",        2, 32, 32, 128, 1, 256, 3, 1, false, false, true, false, false},
      {"slab 3x3 128->64 @64x64 B1 plain",                    1, 64, 64, 128, 1, 64, 3, 1, false, false, false, false, false},
  };
  for (const BwdCase& c : slab_cases) run_dgrad(c);
  g_tiled = 1;
  for (const BwdCase& c : slab_cases) run_dgrad(c);
  g_tiled = 0;
  // data gradients through the CTA-pair tile kernel (full grids, 128 channels), then through the single-CTA kernel
  const BwdCase pair_cases[] = {
      {"pair 3x3 128->128 @32x32 B16 mask+add+add2 inplace",  16, 32, 32, 128, 1, 128, 3, 1, false, true, true, true, true},
      {"pair 3x3 [128|128]->128 @32x32 B16 cat mask+add2",    16, 32, 32, 128, 2, 128, 3, 1, false, true, false, true, false},
      {"pair 1x1 128->128 @64x64 B4 mask",                    4, 64, 64, 128, 1, 128, 1, 1, false, true, false, false, false},
      {"pair 3x3 64->64 @64x64 B5 mask+add2 (BN64)",          5, 64, 64, 64, 1, 64, 3, 1, false, true, false, true, false},
  };
  {
    const char* slab_was = getenv("VUB_ENABLE_SLAB");
    const std::string slab_keep = slab_was ? slab_was : "";
    setenv("VUB_ENABLE_SLAB", "0", 1);
    for (const BwdCase& c : pair_cases) run_dgrad(c);
    setenv("VUB_PAIR", "0", 1);
    for (const BwdCase& c : pair_cases) run_dgrad(c);
    unsetenv("VUB_PAIR");
    if (slab_was) setenv("VUB_ENABLE_SLAB", slab_keep.c_str(), 1);
    else unsetenv("VUB_ENABLE_SLAB");
  }
  // shapes the row-streaming kernel takes over in data-gradient mode (width >= 128, <= 64 channels)
  const BwdCase rows_cases[] = {
      {"rows 1x1 32->32 @8x256 B2 mask+addend inplace",       2, 8, 256, 32, 1, 32, 1, 1, false, true, true, false, true},
      {"rows 3x3 64->64 @8x128 B2 mask+add+add2 inplace",     2, 8, 128, 64, 1, 64, 3, 1, false, true, true, true, true},
      {"rows 3x3 [64|64]->64 @16x128 B1 cat (two launches)",  1, 16, 128, 64, 2, 64, 3, 1, false, true, false, true, false},
      {"rows 1x1 64->64 @4x128 B3 mask",                      3, 4, 128, 64, 1, 64, 1, 1, false, true, false, false, false},
      {"rows 3x3 32->3(32) out_conv @32x128 B2",              2, 32, 128, 32, 1, 32, 3, 1, false, false, false, false, false},
  };
  for (const BwdCase& c : rows_cases) run_dgrad(c);
  // weight gradient of the wide 32-channel layers: pixel splits crossing images / column strips
  {
    const BwdCase w1 = {"wgrad wide 3x3 32->32 @64x256 B3", 3, 64, 256, 32, 1, 32, 3, 1, false, false, false, false, false};
    const BwdCase w2 = {"wgrad wide 3x3 [32|32]->32 @32x128 B5", 5, 32, 128, 32, 2, 32, 3, 1, false, false, false, false, false};
    const BwdCase w3 = {"wgrad wide 3x3 32->3(32) @2x128 B7 (2-row images)", 7, 2, 128, 32, 1, 32, 3, 1, false, false, false, false, false};
    run_wgrad(w1);
    run_wgrad(w2);
    run_wgrad(w3);
  }
  setenv("VUB_ROWS_GRID", "5", 1);  // few persistent CTAs: row ranges cross image boundaries, ring wraps
  {
    const BwdCase c1 = {"rows grid=5: 3x3 [32|32]->32 @32x128 B5 cat mask+add2", 5, 32, 128, 32, 2, 32, 3, 1, false, true, false, true, false};
    run_dgrad(c1);
  }
  unsetenv("VUB_ROWS_GRID");
  {
    const BwdCase big = {"3x3 64->64 @128x128 B2 (many pixel tiles per CTA)", 2, 128, 128, 64, 1, 64, 3, 1, false, true, false, true, false};
    run_wgrad(big);
    const BwdCase big2 = {"3x3 32->32 @64x256 B2", 2, 64, 256, 32, 1, 32, 3, 1, false, true, false, true, false};
    run_wgrad(big2);
    run_dgrad(big2);
  }
  test_wn_prep_wd();
  test_wn_bwd();
  test_stem_bwd(2, 3, 32, 64);
  test_stem_bwd(3, 30, 16, 16);
  test_latent_bwd(3, 128, 4, 4, true, true);
  test_latent_bwd(2, 128, 8, 8, true, false);
  test_latent_bwd(2, 128, 8, 8, false, true);
  test_pad_convert(2, 3, 16, 32, 32);
  if (perf) {
    const int B = 16;
    perf_bwd("32->32 3x3 @256^2", B, 256, 256, 32, 1, 32, 3, 1, false);
    perf_bwd("[32|32]->32 3x3 @256^2", B, 256, 256, 32, 2, 32, 3, 1, false);
    perf_bwd("64->64 3x3 @128^2", B, 128, 128, 64, 1, 64, 3, 1, false);
    perf_bwd("128->128 3x3 @64^2", B, 64, 64, 128, 1, 128, 3, 1, false);
    perf_bwd("[128|128]->128 3x3 @64^2", B, 64, 64, 128, 2, 128, 3, 1, false);
    perf_bwd("128->256 3x3 D2S @64^2", B, 64, 64, 128, 1, 256, 3, 1, true);
    perf_bwd("32->64 3x3 s2 @256^2", B, 256, 256, 32, 1, 64, 3, 2, false);
    perf_bwd("32->32 1x1 @256^2", B, 256, 256, 32, 1, 32, 1, 1, false);
  }
  printf("%s: %d failed\n", g_fail ? "FAILED" : "ALL PASSED", g_fail);
  return g_fail;
}
