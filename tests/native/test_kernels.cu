// Native GPU test + micro-benchmark harness for libvunet_b200.so (runs on the B200 box via gpurun;
// no Python / torch needed).  Every case compares the CUDA path with a straightforward CPU
// computation on the same bf16-rounded inputs.  Exit code = number of failed cases.
//
//   ./test_kernels            -> all correctness cases
//   ./test_kernels perf       -> plus timing of the dominant H256 layer shapes
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

static uint32_t g_seed = 12345u;
static float frand() {  // uniform in [-1, 1)
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFFFF) / 8388608.0f - 1.0f;
}
static float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
static float elu_ref(float v) { return v > 0.f ? v : expm1f(v); }

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  explicit DevBuf(size_t n_) : n(n_) { CK(cudaMalloc(&p, n * sizeof(T))); CK(cudaMemset(p, 0xFF, n * sizeof(T))); }
  ~DevBuf() { cudaFree(p); }
  void up(const std::vector<T>& h) { CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
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

struct Cmp {
  double num = 0, den = 0, maxabs = 0;
  size_t bad = 0, n = 0;
  void add(float got, float ref, float atol, float rtol) {
    const double d = (double)got - ref;
    num += d * d;
    den += (double)ref * ref;
    if (fabs(d) > maxabs) maxabs = fabs(d);
    if (!(fabs(d) <= atol + rtol * fabs(ref))) ++bad;  // also catches NaN
    ++n;
  }
  double rel() const { return sqrt(num / (den > 0 ? den : 1)); }
};

static int g_fail = 0;
static void report(const std::string& name, const Cmp& c) {
  const bool ok = c.bad == 0 && c.n > 0;
  printf("[%s] %-58s n=%zu relL2=%.3e maxabs=%.3e bad=%zu\n", ok ? "PASS" : "FAIL", name.c_str(), c.n,
         c.rel(), c.maxabs, c.bad);
  if (!ok) ++g_fail;
}

// ------------------------------------------------------------------------------------------------
struct ConvCase {
  const char* name;
  int B, H, W, c_src, nsrc, cout, cout_valid, ksize, stride;
  bool residual, raw, elu, d2s, f32nhwc, f32nchw;
  int stages;
};


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

static int g_elu_in = 0;  // VUB_CONV_ELU_IN source mask for the next run_conv_case calls

// VUB_SANITIZE=1 (tools/sanitize.sh): compute-sanitizer slows kernels down 10-100x; keep the cases below 16 k pixels
// (every kernel family still runs, incl. the row, tile, CTA-pair, slab and per-tap variants)
static bool sanitize_skip(long long pixels) {
  static const bool on = getenv("VUB_SANITIZE") != nullptr;
  return on && pixels > 16384;
}

static void run_conv_case(const ConvCase& cc) {
  if (sanitize_skip((long long)cc.B * cc.H * cc.W)) return;
  const int taps = cc.ksize * cc.ksize;
  const int K = cc.nsrc * taps * cc.c_src;
  const int Ho = cc.H / cc.stride, Wo = cc.W / cc.stride;
  const size_t in_elems = (size_t)cc.B * cc.H * cc.W * cc.c_src;
  std::vector<float> x[2];
  for (int s = 0; s < cc.nsrc; ++s) {
    x[s].resize(in_elems);
    for (auto& v : x[s]) v = bf16r(frand());
  }
  // ELU-on-load: the device gets the raw values, the reference convolves bf16(ELU(raw))
  std::vector<float> xraw[2] = {x[0], x[1]};
  for (int s = 0; s < cc.nsrc; ++s)
    if ((g_elu_in >> s) & 1)
      for (auto& v : x[s]) v = bf16r(v > 0.f ? v : expm1f(v));
  std::vector<float> w((size_t)cc.cout * K, 0.f);
  for (int co = 0; co < cc.cout_valid; ++co)
    for (int k = 0; k < K; ++k) w[(size_t)co * K + k] = bf16r(frand() * 0.08f);
  std::vector<float> scale(cc.cout), shift(cc.cout);
  for (int i = 0; i < cc.cout; ++i) {
    scale[i] = 0.7f + 0.6f * fabsf(frand());
    shift[i] = 0.2f * frand();
  }
  const size_t out_elems = (size_t)cc.B * Ho * Wo * cc.cout;
  std::vector<float> res;
  if (cc.residual) {
    res.resize(out_elems);
    for (auto& v : res) v = bf16r(frand());
  }
  // CPU reference, NHWC [B,Ho,Wo,cout]
  std::vector<float> ref(out_elems);
  const int pad = cc.ksize / 2;
  for (int n = 0; n < cc.B; ++n)
    for (int oh = 0; oh < Ho; ++oh)
      for (int ow = 0; ow < Wo; ++ow) {
        float* o = &ref[(((size_t)n * Ho + oh) * Wo + ow) * cc.cout];
        std::vector<double> acc(cc.cout, 0.0);
        for (int s = 0; s < cc.nsrc; ++s)
          for (int t = 0; t < taps; ++t) {
            const int r = t / cc.ksize, q = t % cc.ksize;
            const int ih = oh * cc.stride + r - pad, iw = ow * cc.stride + q - pad;
            if (ih < 0 || ih >= cc.H || iw < 0 || iw >= cc.W) continue;
            const float* xp = &x[s][(((size_t)n * cc.H + ih) * cc.W + iw) * cc.c_src];
            for (int co = 0; co < cc.cout_valid; ++co) {
              const float* wp = &w[(size_t)co * K + (size_t)(s * taps + t) * cc.c_src];
              float a = 0.f;
              for (int c = 0; c < cc.c_src; ++c) a += xp[c] * wp[c];
              acc[co] += a;
            }
          }
        for (int co = 0; co < cc.cout; ++co) {
          float y = scale[co] * (float)acc[co] + shift[co];
          if (cc.residual) y += res[(((size_t)n * Ho + oh) * Wo + ow) * cc.cout + co];
          o[co] = y;
        }
      }

  DevBuf<__nv_bfloat16> dx0(in_elems), dx1(cc.nsrc == 2 ? in_elems : 8), dw(w.size());
  DevBuf<float> dscale(cc.cout), dshift(cc.cout);
  DevBuf<__nv_bfloat16> dres(cc.residual ? out_elems : 8), draw(out_elems), delu(out_elems);
  DevBuf<float> df32(out_elems);
  dx0.up(to_bf16(xraw[0]));
  if (cc.nsrc == 2) dx1.up(to_bf16(xraw[1]));
  dw.up(to_bf16(w));
  DevBuf<__nv_bfloat16> dwt(w.size());
  const bool tiled = g_tiled && cc.c_src % 64 == 0;
  if (tiled) {
    dwt.up(to_bf16(tile_major(w, cc.cout, K)));
    if (getenv("VUB_TEST_ZERO_W")) CK(cudaMemset(dw.p, 0, w.size() * 2));  // prove the kernels read the tiled copy
  }
  dscale.up(scale);
  dshift.up(shift);
  if (cc.residual) dres.up(to_bf16(res));

  vub_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.x[0] = dx0.p;
  d.x[1] = cc.nsrc == 2 ? dx1.p : nullptr;
  d.w = dw.p;
  d.scale = dscale.p;
  d.shift = dshift.p;
  d.residual = cc.residual ? dres.p : nullptr;
  const bool f32 = cc.f32nhwc || cc.f32nchw;
  d.out_raw = (!f32 && cc.raw) ? draw.p : nullptr;
  d.out_elu = (!f32 && cc.elu) ? delu.p : nullptr;
  d.out_f32 = f32 ? df32.p : nullptr;
  d.nsrc = cc.nsrc;
  d.c_src = cc.c_src;
  d.batch = cc.B;
  d.hin = cc.H;
  d.win = cc.W;
  d.cout = cc.cout;
  d.cout_valid = cc.cout_valid;
  d.ksize = cc.ksize;
  d.stride = cc.stride;
  d.flags = (cc.d2s ? VUB_CONV_D2S : 0) | (cc.f32nhwc ? VUB_CONV_F32_NHWC : 0) | (cc.f32nchw ? VUB_CONV_F32_NCHW : 0) |
            ((g_elu_in & 1) ? VUB_CONV_ELU_IN0 : 0) | ((g_elu_in & 2) ? VUB_CONV_ELU_IN1 : 0);
  d.stages = cc.stages;
  d.w_tiled = tiled ? dwt.p : nullptr;
  int rc = vub_conv_fwd(&d, nullptr);
  if (rc != 0) {
    printf("[FAIL] %-58s launch rc=%d (%s)\n", cc.name, rc, vub_last_error());
    ++g_fail;
    return;
  }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    printf("[FAIL] %-58s sync error %s\n", cc.name, cudaGetErrorString(se));
    ++g_fail;
    exit(101);  // context is dead
  }

  auto ref_at = [&](int n, int oh, int ow, int co) { return ref[(((size_t)n * Ho + oh) * Wo + ow) * cc.cout + co]; };
  if (f32) {
    std::vector<float> got = df32.down();
    Cmp c;
    for (int n = 0; n < cc.B; ++n)
      for (int oh = 0; oh < Ho; ++oh)
        for (int ow = 0; ow < Wo; ++ow)
          for (int co = 0; co < cc.cout_valid; ++co) {
            const size_t gi = cc.f32nchw ? (((size_t)n * cc.cout_valid + co) * Ho + oh) * Wo + ow
                                         : (((size_t)n * Ho + oh) * Wo + ow) * cc.cout + co;
            c.add(got[gi], ref_at(n, oh, ow, co), 2e-3f, 2e-3f);
          }
    report(std::string(cc.name) + " f32", c);
    return;
  }
  for (int which = 0; which < 2; ++which) {
    if ((which == 0 && !cc.raw) || (which == 1 && !cc.elu)) continue;
    std::vector<__nv_bfloat16> got = which == 0 ? draw.down() : delu.down();
    Cmp c;
    for (int n = 0; n < cc.B; ++n)
      for (int oh = 0; oh < Ho; ++oh)
        for (int ow = 0; ow < Wo; ++ow)
          for (int co = 0; co < cc.cout; ++co) {
            size_t gi;
            if (cc.d2s) {
              const int cq = cc.cout / 4, qd = co / cq, cp = co % cq;
              gi = (((size_t)n * 2 * Ho + 2 * oh + (qd >> 1)) * 2 * Wo + 2 * ow + (qd & 1)) * cq + cp;
            } else {
              gi = (((size_t)n * Ho + oh) * Wo + ow) * cc.cout + co;
            }
            float r = ref_at(n, oh, ow, co);
            if (which == 1) r = elu_ref(r);
            c.add(__bfloat162float(got[gi]), r, 6e-3f, 1e-2f);
          }
    report(std::string(cc.name) + (which == 0 ? " raw" : " elu"), c);
  }
}

// ------------------------------------------------------------------------------------------------
static void test_wn_prep() {
  struct L { int cout, cin, k, c_src; };
  const L layers[] = {{32, 32, 3, 32}, {128, 256, 1, 128}, {64, 128, 3, 64}, {3, 32, 3, 32}, {128, 128, 3, 128}};
  const int nl = 5;
  std::vector<vub_wn_entry> table(nl);
  std::vector<std::vector<float>> v(nl), g(nl), ga(nl), be(nl), bi(nl);
  std::vector<DevBuf<float>*> dv, dg, dga, dbe, dbi, dsc, dsh;
  std::vector<DevBuf<__nv_bfloat16>*> dw;
  int rows = 0;
  for (int l = 0; l < nl; ++l) {
    const L& L_ = layers[l];
    const int len = L_.cin * L_.k * L_.k;
    v[l].resize((size_t)L_.cout * len);
    for (auto& t : v[l]) t = frand();
    g[l].resize(L_.cout); ga[l].resize(L_.cout); be[l].resize(L_.cout); bi[l].resize(L_.cout);
    for (int i = 0; i < L_.cout; ++i) { g[l][i] = 0.5f + fabsf(frand()); ga[l][i] = 0.7f + 0.6f * fabsf(frand()); be[l][i] = 0.1f * frand(); bi[l][i] = 0.1f * frand(); }
    dv.push_back(new DevBuf<float>(v[l].size())); dv[l]->up(v[l]);
    dg.push_back(new DevBuf<float>(L_.cout)); dg[l]->up(g[l]);
    dga.push_back(new DevBuf<float>(L_.cout)); dga[l]->up(ga[l]);
    dbe.push_back(new DevBuf<float>(L_.cout)); dbe[l]->up(be[l]);
    dbi.push_back(new DevBuf<float>(L_.cout)); dbi[l]->up(bi[l]);
    dsc.push_back(new DevBuf<float>(L_.cout)); dsh.push_back(new DevBuf<float>(L_.cout));
    dw.push_back(new DevBuf<__nv_bfloat16>((size_t)L_.cout * len));
    vub_wn_entry& e = table[l];
    e.v = dv[l]->p; e.g = dg[l]->p; e.gamma = dga[l]->p; e.beta = dbe[l]->p; e.bias = dbi[l]->p;
    e.w_packed = dw[l]->p; e.scale = dsc[l]->p; e.shift = dsh[l]->p;
    e.cout = L_.cout; e.cin = L_.cin; e.ktaps = L_.k * L_.k; e.c_src = L_.c_src; e.ktot = len; e.row_start = rows;
    rows += L_.cout;
  }
  vub_wn_entry* dtab;
  CK(cudaMalloc(&dtab, sizeof(vub_wn_entry) * nl));
  CK(cudaMemcpy(dtab, table.data(), sizeof(vub_wn_entry) * nl, cudaMemcpyHostToDevice));
  int rc = vub_wn_prep(dtab, nl, rows, nullptr);
  if (rc) { printf("[FAIL] wn_prep rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  CK(cudaDeviceSynchronize());
  Cmp cw, cs;
  for (int l = 0; l < nl; ++l) {
    const L& L_ = layers[l];
    const int taps = L_.k * L_.k, len = L_.cin * taps;
    auto w = dw[l]->down();
    auto sc = dsc[l]->down();
    auto sh = dsh[l]->down();
    for (int co = 0; co < L_.cout; ++co) {
      double ss = 0;
      for (int i = 0; i < len; ++i) ss += (double)v[l][(size_t)co * len + i] * v[l][(size_t)co * len + i];
      const double inv = 1.0 / sqrt(ss);
      for (int ci = 0; ci < L_.cin; ++ci)
        for (int t = 0; t < taps; ++t) {
          const int src = ci / L_.c_src, cs_ = ci % L_.c_src;
          const float got = __bfloat162float(w[(size_t)co * len + (src * taps + t) * L_.c_src + cs_]);
          cw.add(got, (float)(v[l][((size_t)co * L_.cin + ci) * taps + t] * inv), 1e-4f, 8e-3f);
        }
      cs.add(sc[co], ga[l][co] * g[l][co], 1e-6f, 1e-6f);
      cs.add(sh[co], ga[l][co] * bi[l][co] + be[l][co], 1e-6f, 1e-6f);
    }
  }
  report("wn_prep packed weights (5 layers, split sources)", cw);
  report("wn_prep scale/shift", cs);
  cudaFree(dtab);
}

static void test_stem(int B, int cx, int H, int W) {
  const int cout = 32;
  std::vector<float> x((size_t)B * cx * H * W), v((size_t)cout * cx), g(cout), ga(cout), be(cout), bi(cout);
  for (auto& t : x) t = frand();
  for (auto& t : v) t = frand();
  for (int i = 0; i < cout; ++i) { g[i] = 0.5f + fabsf(frand()); ga[i] = 0.7f + 0.6f * fabsf(frand()); be[i] = 0.1f * frand(); bi[i] = 0.1f * frand(); }
  DevBuf<float> dx(x.size()), dv(v.size()), dg(cout), dga(cout), dbe(cout), dbi(cout);
  dx.up(x); dv.up(v); dg.up(g); dga.up(ga); dbe.up(be); dbi.up(bi);
  const size_t oe = (size_t)B * H * W * cout;
  DevBuf<__nv_bfloat16> draw(oe), delu(oe);
  int rc = vub_stem_fwd(dx.p, dv.p, dg.p, dga.p, dbe.p, dbi.p, B, cx, H, W, cout, draw.p, delu.p, nullptr);
  if (rc) { printf("[FAIL] stem rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  CK(cudaDeviceSynchronize());
  auto raw = draw.down();
  auto el = delu.down();
  Cmp c0, c1;
  for (int n = 0; n < B; ++n)
    for (int p = 0; p < H * W; ++p)
      for (int co = 0; co < cout; ++co) {
        double ss = 0, acc = 0;
        for (int ci = 0; ci < cx; ++ci) {
          ss += (double)v[co * cx + ci] * v[co * cx + ci];
          acc += (double)v[co * cx + ci] * x[((size_t)n * cx + ci) * H * W + p];
        }
        const float y = (float)(ga[co] * (g[co] * acc / sqrt(ss) + bi[co]) + be[co]);
        const size_t gi = ((size_t)n * H * W + p) * cout + co;
        c0.add(__bfloat162float(raw[gi]), y, 4e-3f, 8e-3f);
        c1.add(__bfloat162float(el[gi]), elu_ref(y), 4e-3f, 8e-3f);
      }
  char nm[96];
  snprintf(nm, sizeof nm, "stem B%d cx%d %dx%d raw", B, cx, H, W);
  report(nm, c0);
  snprintf(nm, sizeof nm, "stem B%d cx%d %dx%d elu", B, cx, H, W);
  report(nm, c1);
}

static void test_latent(int B, int C, int h, int w) {
  const size_t n = (size_t)B * C * h * w;
  std::vector<float> ml((size_t)B * h * w * 2 * C), eps(n);
  for (auto& t : ml) t = 2.f * frand();
  for (auto& t : eps) t = 1.5f * frand();
  DevBuf<float> dml(ml.size()), deps(n), dmu(n), dls(n), dz(n);
  DevBuf<__nv_bfloat16> dez(n), dem(n);
  const int np = vub_latent_num_partials(B, C, h, w);
  DevBuf<float> dkl(np);
  dml.up(ml); deps.up(eps);
  int rc = vub_latent_fwd(dml.p, deps.p, B, C, h, w, dmu.p, dls.p, dz.p, dez.p, dem.p, dkl.p, nullptr);
  if (rc) { printf("[FAIL] latent rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  CK(cudaDeviceSynchronize());
  auto mu = dmu.down(); auto ls = dls.down(); auto z = dz.down(); auto ez = dez.down(); auto em = dem.down(); auto kl = dkl.down();
  Cmp c, cb, ck;
  double klref = 0, klgot = 0;
  for (int b = 0; b < B; ++b)
    for (int ch = 0; ch < C; ++ch)
      for (int p = 0; p < h * w; ++p) {
        const float m = ml[((size_t)b * h * w + p) * 2 * C + ch];
        const float lg = ml[((size_t)b * h * w + p) * 2 * C + C + ch];
        const float l = 1.f / (1.f + expf(-lg));
        const float sd = expf(l);
        const size_t o = ((size_t)b * C + ch) * h * w + p;
        const float zz = eps[o] * sd + m;
        c.add(mu[o], m, 1e-6f, 1e-6f);
        c.add(ls[o], l, 1e-6f, 1e-5f);
        c.add(z[o], zz, 1e-5f, 1e-5f);
        const size_t oh = ((size_t)b * h * w + p) * C + ch;
        cb.add(__bfloat162float(ez[oh]), elu_ref(zz), 4e-3f, 8e-3f);
        cb.add(__bfloat162float(em[oh]), elu_ref(m), 4e-3f, 8e-3f);
        klref += -l + 0.5 * ((double)sd * sd + (double)m * m);
      }
  for (float t : kl) klgot += t;
  ck.add((float)klgot, (float)klref, 1e-3f, 1e-4f);
  report("latent mu/logstd/z (fp32 NCHW)", c);
  report("latent elu(z)/elu(mu) (bf16 NHWC)", cb);
  report("latent KL partial sums", ck);
}

static void test_convert(int B, int C, int H, int W) {
  const size_t n = (size_t)B * C * H * W;
  std::vector<float> x(n);
  for (auto& t : x) t = 2.f * frand();
  DevBuf<float> dx(n), dback(n);
  DevBuf<__nv_bfloat16> draw(n), delu(n);
  dx.up(x);
  int rc = vub_nchw_f32_to_nhwc_bf16(dx.p, B, C, H, W, draw.p, delu.p, nullptr);
  rc |= vub_nhwc_bf16_to_nchw_f32(draw.p, B, C, H, W, dback.p, nullptr);
  if (rc) { printf("[FAIL] convert rc=%d %s\n", rc, vub_last_error()); ++g_fail; return; }
  CK(cudaDeviceSynchronize());
  auto raw = draw.down(); auto el = delu.down(); auto back = dback.down();
  Cmp c;
  for (int b = 0; b < B; ++b)
    for (int ch = 0; ch < C; ++ch)
      for (int p = 0; p < H * W; ++p) {
        const float v = x[((size_t)b * C + ch) * H * W + p];
        const size_t o = ((size_t)b * H * W + p) * C + ch;
        c.add(__bfloat162float(raw[o]), bf16r(v), 0.f, 0.f);
        c.add(__bfloat162float(el[o]), bf16r(elu_ref(v)), 1e-6f, 8e-3f);
        c.add(back[((size_t)b * C + ch) * H * W + p], bf16r(v), 0.f, 0.f);
      }
  char nm[96];
  snprintf(nm, sizeof nm, "layout converters B%d C%d %dx%d", B, C, H, W);
  report(nm, c);
}

// ------------------------------------------------------------------------------------------------
static void perf_conv(const char* name, int B, int H, int W, int c_src, int nsrc, int cout, int ksize,
                      int stride, bool residual, bool dual, bool d2s, int stages) {
  const int taps = ksize * ksize, K = nsrc * taps * c_src;
  const int Ho = H / stride, Wo = W / stride;
  const size_t in_elems = (size_t)B * H * W * c_src, out_elems = (size_t)B * Ho * Wo * cout;
  DevBuf<__nv_bfloat16> dx0(in_elems), dx1(nsrc == 2 ? in_elems : 8), dw((size_t)cout * K), dres(residual ? out_elems : 8),
      draw(out_elems), delu(dual ? out_elems : 8);
  DevBuf<float> dscale(cout), dshift(cout);
  CK(cudaMemset(dx0.p, 0, in_elems * 2));
  if (nsrc == 2) CK(cudaMemset(dx1.p, 0, in_elems * 2));
  CK(cudaMemset(dw.p, 0, (size_t)cout * K * 2));
  CK(cudaMemset(dscale.p, 0, cout * 4));
  CK(cudaMemset(dshift.p, 0, cout * 4));
  if (residual) CK(cudaMemset(dres.p, 0, out_elems * 2));
  vub_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.x[0] = dx0.p; d.x[1] = nsrc == 2 ? dx1.p : nullptr; d.w = dw.p; d.scale = dscale.p; d.shift = dshift.p;
  d.residual = residual ? dres.p : nullptr; d.out_raw = draw.p; d.out_elu = dual ? delu.p : nullptr;
  d.nsrc = nsrc; d.c_src = c_src; d.batch = B; d.hin = H; d.win = W; d.cout = cout; d.cout_valid = cout;
  d.ksize = ksize; d.stride = stride; d.flags = d2s ? VUB_CONV_D2S : 0; d.stages = stages;
  DevBuf<__nv_bfloat16> dwt((size_t)cout * K);
  if (getenv("VUB_TEST_TILED") && c_src % 64 == 0) {
    CK(cudaMemset(dwt.p, 0, (size_t)cout * K * 2));
    d.w_tiled = dwt.p;
  }
  if (g_elu_in) {  // lean I/O: raw source(s) transformed on load; the residual IS the raw source (an L2 hit)
    d.flags |= ((g_elu_in & 1) ? VUB_CONV_ELU_IN0 : 0) | ((g_elu_in & 2) ? VUB_CONV_ELU_IN1 : 0);
    if (residual) d.residual = dx0.p;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) vub_conv_fwd(&d, nullptr);
  CK(cudaDeviceSynchronize());
  const int iters = 20;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) vub_conv_fwd(&d, nullptr);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  const double flop = 2.0 * B * Ho * Wo * (double)cout * K;
  const double bytes = 2.0 * (nsrc * (double)in_elems + out_elems * (1 + (dual ? 1 : 0) + (residual && !g_elu_in ? 1 : 0)) + (double)cout * K);
  printf("[PERF] %-44s st=%d  %8.3f ms  %7.1f TFLOP/s  %7.1f GB/s(algorithmic)\n", name, stages, ms,
         flop / ms * 1e-9, bytes / ms * 1e-6);
}

int main(int argc, char** argv) {
  const bool perf = argc > 1 && !strcmp(argv[1], "perf");
  if (!perf && !getenv("VUB_ENABLE_SLAB")) setenv("VUB_ENABLE_SLAB", "1", 1);  // correctness runs cover the opt-in slab kernel
  if (getenv("VUB_TEST_ELU_IN")) g_elu_in = atoi(getenv("VUB_TEST_ELU_IN"));
  if (argc >= 14 && !strcmp(argv[1], "layer")) {
    // layer B H W c_src nsrc cout k stride res dual d2s stages  -> time one layer shape (for ncu captures)
    int a[12];
    for (int i = 0; i < 12; ++i) a[i] = atoi(argv[2 + i]);
    perf_conv("layer", a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8] != 0, a[9] != 0, a[10] != 0, a[11]);
    return 0;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s  sm_%d%d  SMs=%d  abi=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         vub_abi_version());

  // Smallest maps (<= 8x8): split-K over a thread-block cluster (partials reduced through distributed shared memory),
  // and the 256 / 512-channel sources of VunetOrg's 4-group prior (models/vunets.py:655-676).
  const ConvCase small_cases[] = {
      {"3x3 128->128 @4x4 B5 +res dual (split 3)",         5, 4, 4, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 128->128 @8x8 B9 +res dual (split 3, 5 tiles)", 9, 8, 8, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 [128|128]->128 @8x8 B3 cat +res dual (split 4)", 3, 8, 8, 128, 2, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 [64|64]->64 @4x4 B3 cat +res dual (split 3)",  3, 4, 4, 64, 2, 64, 64, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 128->512 D2S @4x4 B3 dual (split 3)",          3, 4, 4, 128, 1, 512, 512, 3, 1, false, true, true, true, false, false, 0},
      {"3x3 128->256 f32 NHWC @8x8 B3 (latent, split 3)",  3, 8, 8, 128, 1, 256, 256, 3, 1, false, false, false, false, true, false, 0},
      {"3x3 s2 128->128 @8x8 B3 -> 4x4 (split 3)",         3, 8, 8, 128, 1, 128, 128, 3, 2, false, true, true, false, false, false, 0},
      {"3x3 512->128 f32 NHWC @2x2 B5 (auto_lp, split 8)", 5, 2, 2, 512, 1, 128, 128, 3, 1, false, false, false, false, true, false, 0},
      {"3x3 512->128 f32 NHWC @4x4 B9 (auto_lp, split 8)", 9, 4, 4, 512, 1, 128, 128, 3, 1, false, false, false, false, true, false, 0},
      {"3x3 [512|512]->512 @4x4 B3 cat +res dual (split 8)", 3, 4, 4, 512, 2, 512, 512, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 [512|512]->512 @2x2 B33 cat +res raw (split 8)", 33, 2, 2, 512, 2, 512, 512, 3, 1, true, true, false, false, false, false, 0},
      {"3x3 [256|256]->256 @1x1 B3 cat +res dual (split 8)", 3, 1, 1, 256, 2, 256, 256, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 256->64 f32 NHWC @2x2 B2 (auto_lp, split 4)",  2, 2, 2, 256, 1, 64, 64, 3, 1, false, false, false, false, true, false, 0},
      {"1x1 128->512 @4x4 B3 elu only",                    3, 4, 4, 128, 1, 512, 512, 1, 1, false, false, true, false, false, false, 0},
      {"1x1 [128|128]->128 @8x8 B3 cat dual (latent_nins)", 3, 8, 8, 128, 2, 128, 128, 1, 1, false, true, true, false, false, false, 0},
      {"3x3 512->128 f32 NHWC @16x16 B2 (wide, no split)", 2, 16, 16, 512, 1, 128, 128, 3, 1, false, false, false, false, true, false, 0},
      // linear layers of the behavior net (nn.Linear / LSTMCell rows as a 1x1 conv on a 1x1 map): reduce-scatter split-K
      {"linear 2048->2048 B200 elu only (reduce-scatter)",  200, 1, 1, 2048, 1, 2048, 2048, 1, 1, false, false, true, false, false, false, 0},
      {"linear 1088->4096 B256 f32 (LSTM gates)",           256, 1, 1, 1088, 1, 4096, 4096, 1, 1, false, false, false, false, true, false, 0},
      {"linear 2048->512 B130 f32 (last MLP layer)",        130, 1, 1, 2048, 1, 512, 512, 1, 1, false, false, false, false, true, false, 0},
      {"linear 512->2048 B256 dual +res (no split: 8 k-steps)", 256, 1, 1, 512, 1, 2048, 2048, 1, 1, true, true, true, false, false, false, 0},
      {"linear 1024->1024 B64 raw",                          64, 1, 1, 1024, 1, 1024, 1024, 1, 1, false, true, false, false, false, false, 0},
  };
  if (argc > 1 && !strcmp(argv[1], "splitk")) {
    // 1: the library's rule (three k-ranges on maps up to 4x4); 0: off; n >= 2: n ranges on every map up to 8x8
    for (const char* mode : {"1", "0", "2", "4", "8"}) {
      setenv("VUB_SPLITK", mode, 1);
      printf("--- correctness, VUB_SPLITK=%s\n", mode);
      for (const ConvCase& c : small_cases) run_conv_case(c);
    }
    unsetenv("VUB_SPLITK");
    for (const char* rs : {"8", "4", "2", "0"}) {  // reduce-scatter ranges for the linear layers
      setenv("VUB_LINEAR_RS", rs, 1);
      printf("--- correctness + timing, VUB_LINEAR_RS=%s\n", rs);
      for (const ConvCase& c : small_cases)
        if (!strncmp(c.name, "linear", 6)) run_conv_case(c);
      perf_conv("linear 2048->2048 B256 elu", 256, 1, 1, 2048, 1, 2048, 1, 1, false, false, false, 0);
      perf_conv("linear 1088->4096 B256", 256, 1, 1, 1088, 1, 4096, 1, 1, false, false, false, 0);
      perf_conv("linear 1024->1024 B256", 256, 1, 1, 1024, 1, 1024, 1, 1, false, false, false, 0);
    }
    unsetenv("VUB_LINEAR_RS");
    for (int sk : {1, 0, 2}) {
      char buf[4];
      snprintf(buf, sizeof buf, "%d", sk);
      setenv("VUB_SPLITK", buf, 1);
      printf("--- VUB_SPLITK=%d\n", sk);
      perf_conv("128->128 3x3 @4^2 B64 +res dual", 64, 4, 4, 128, 1, 128, 3, 1, true, true, false, 0);
      perf_conv("128->128 3x3 @8^2 B64 +res dual", 64, 8, 8, 128, 1, 128, 3, 1, true, true, false, 0);
      perf_conv("[128|128]->128 3x3 @4^2 B64 +res dual", 64, 4, 4, 128, 2, 128, 3, 1, true, true, false, 0);
      perf_conv("[128|128]->128 3x3 @8^2 B64 +res dual", 64, 8, 8, 128, 2, 128, 3, 1, true, true, false, 0);
      perf_conv("128->512 3x3 D2S @8^2 B64 dual", 64, 8, 8, 128, 1, 512, 3, 1, false, true, true, 0);
      perf_conv("128->128 3x3 @8^2 B8 +res dual", 8, 8, 8, 128, 1, 128, 3, 1, true, true, false, 0);
      perf_conv("[128|128]->128 3x3 @8^2 B8 +res dual", 8, 8, 8, 128, 2, 128, 3, 1, true, true, false, 0);
      perf_conv("512->128 3x3 @4^2 B64 (auto_lp)", 64, 4, 4, 512, 1, 128, 3, 1, false, false, false, 0);
      perf_conv("[512|512]->512 3x3 @4^2 B64 +res dual", 64, 4, 4, 512, 2, 512, 3, 1, true, true, false, 0);
      perf_conv("[512|512]->512 3x3 @2^2 B64 +res dual", 64, 2, 2, 512, 2, 512, 3, 1, true, true, false, 0);
    }
    unsetenv("VUB_SPLITK");
    printf("%s: %d failed\n", g_fail ? "FAILED" : "ALL PASSED", g_fail);
    return g_fail;
  }
  for (const ConvCase& c : small_cases) run_conv_case(c);

  //                name                                   B  H   W  c  ns co  cv k s  res  raw  elu  d2s  f32h f32c st
  const ConvCase cases[] = {
      {"3x3 128->128 @16x16 B2 (BN128,BKC64)",             2, 16, 16, 128, 1, 128, 128, 3, 1, false, true, false, false, false, false, 0},
      {"3x3 128->128 @16x16 B2 +res dual",                 2, 16, 16, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 32->32 @8x256 B1 (BN32,BKC32) +res dual",      1, 8, 256, 32, 1, 32, 32, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 [32|32]->32 @16x128 B2 cat +res dual",         2, 16, 128, 32, 2, 32, 32, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 64->64 @32x32 B3 (BN64,BKC64) +res dual",      3, 32, 32, 64, 1, 64, 64, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 [64|64]->64 @64x64 B1 cat +res raw",           1, 64, 64, 64, 2, 64, 64, 3, 1, true, true, false, false, false, false, 0},
      {"3x3 [128|128]->128 @8x8 B3 cat +res dual",         3, 8, 8, 128, 2, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 128->128 @4x4 B5 (tile spans 8 images)",       5, 4, 4, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"3x3 128->128 @2x2 B7 (tile spans 32 images)",      7, 2, 2, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"1x1 128->128 @8x8 B2 elu only",                    2, 8, 8, 128, 1, 128, 128, 1, 1, false, false, true, false, false, false, 0},
      {"1x1 [128|128]->128 @4x4 B2 cat elu only",          2, 4, 4, 128, 2, 128, 128, 1, 1, false, false, true, false, false, false, 0},
      {"1x1 32->32 @32x64 B2 elu only",                    2, 32, 64, 32, 1, 32, 32, 1, 1, false, false, true, false, false, false, 0},
      {"1x1 64->64 @32x32 B2 dual",                        2, 32, 32, 64, 1, 64, 64, 1, 1, false, true, true, false, false, false, 0},
      {"3x3 s2 32->64 @64x64 B2 (BN64,BKC32) dual",        2, 64, 64, 32, 1, 64, 64, 3, 2, false, true, true, false, false, false, 0},
      {"3x3 s2 64->128 @32x32 B2 dual",                    2, 32, 32, 64, 1, 128, 128, 3, 2, false, true, true, false, false, false, 0},
      {"3x3 s2 128->128 @8x8 B3 dual",                     3, 8, 8, 128, 1, 128, 128, 3, 2, false, true, true, false, false, false, 0},
      {"3x3 s2 128->128 @4x4 B3 -> 2x2",                   3, 4, 4, 128, 1, 128, 128, 3, 2, false, true, true, false, false, false, 0},
      {"3x3 128->512 D2S @8x8 B2 (2 N tiles) dual",        2, 8, 8, 128, 1, 512, 512, 3, 1, false, true, true, true, false, false, 0},
      {"3x3 128->256 D2S @16x16 B1 dual",                  1, 16, 16, 128, 1, 256, 256, 3, 1, false, true, true, true, false, false, 0},
      {"3x3 64->128 D2S @32x32 B1 dual",                   1, 32, 32, 64, 1, 128, 128, 3, 1, false, true, true, true, false, false, 0},
      {"3x3 128->256 f32 NHWC @4x4 B3 (latent)",           3, 4, 4, 128, 1, 256, 256, 3, 1, false, false, false, false, true, false, 0},
      {"3x3 32->3(16) f32 NCHW @32x128 B2 (out_conv)",     2, 32, 128, 32, 1, 16, 3, 3, 1, false, false, false, false, false, true, 0},
      {"3x3 128->128 @32x32 B2 stages=2",                  2, 32, 32, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 2},
      {"3x3 128->128 @32x32 B2 stages=5",                  2, 32, 32, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 5},
  };
  for (const ConvCase& c : cases) run_conv_case(c);
  g_tiled = 1;  // the same shapes with 1-D bulk weight loads from the tile-major copy (>= 64 channels per source)
  for (const ConvCase& c : cases)
    if (c.c_src % 64 == 0) run_conv_case(c);
  g_tiled = 0;
  // activation-slab kernel (3x3 stride 1, >= 64 channels per source, 64x64 / 32x32)
  const ConvCase slab_cases[] = {
      {"slab 3x3 128->128 @64x64 B2 +res dual",            2, 64, 64, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"slab 3x3 [128|128]->128 @32x32 B3 cat +res dual",  3, 32, 32, 128, 2, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"slab 3x3 64->128 @64x64 B1 raw",                   1, 64, 64, 64, 1, 128, 128, 3, 1, false, true, false, false, false, false, 0},
      {"slab 3x3 128->256 D2S @64x64 B1 dual",             1, 64, 64, 128, 1, 256, 256, 3, 1, false, true, true, true, false, false, 0},
      {"slab 3x3 128->512 D2S @32x32 B2 elu",              2, 32, 32, 128, 1, 512, 512, 3, 1, false, false, true, true, false, false, 0},
      {"slab 3x3 128->128 @32x32 B41 (items > SMs) +res",  41, 32, 32, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
  };
  for (const ConvCase& c : slab_cases) run_conv_case(c);
  g_tiled = 1;
  for (const ConvCase& c : slab_cases) run_conv_case(c);
  g_tiled = 0;
  // CTA-pair tile kernel (conv_pair_kernel): full grids of 128-channel layers, contiguous epilogues; then the same
  // shapes through the single-CTA kernel
  const ConvCase pair_cases[] = {
      {"pair 3x3 128->128 @32x32 B19 (76 pairs > 74) +res", 19, 32, 32, 128, 1, 128, 128, 3, 1, true, true, true, false, false, false, 0},
      {"pair 3x3 [128|128]->128 @32x32 B20 cat +res raw",  20, 32, 32, 128, 2, 128, 128, 3, 1, true, true, false, false, false, false, 0},
      {"pair 3x3 128->256 D2S @32x32 B20 dual (BN256)",    20, 32, 32, 128, 1, 256, 256, 3, 1, false, true, true, true, false, false, 0},
      {"pair 3x3 128->512 D2S @32x32 B10 (2 N tiles) elu", 10, 32, 32, 128, 1, 512, 512, 3, 1, false, false, true, true, false, false, 0},
      {"pair 1x1 128->128 @64x64 B5 elu only",             5, 64, 64, 128, 1, 128, 128, 1, 1, false, false, true, false, false, false, 0},
      {"pair 3x3 [64|64]->64 @128x128 B2 cat +res dual (BN64)", 2, 128, 128, 64, 2, 64, 64, 3, 1, true, true, true, false, false, false, 0},
      {"pair 3x3 64->64 @64x64 B5 +res raw (BN64)",        5, 64, 64, 64, 1, 64, 64, 3, 1, true, true, false, false, false, false, 0},
  };
  {
    const std::string slab_keep = getenv("VUB_ENABLE_SLAB") ? getenv("VUB_ENABLE_SLAB") : "1";
    setenv("VUB_ENABLE_SLAB", "0", 1);
    for (const ConvCase& c : pair_cases) run_conv_case(c);
    setenv("VUB_PAIR", "0", 1);
    for (const ConvCase& c : pair_cases) run_conv_case(c);
    unsetenv("VUB_PAIR");
    setenv("VUB_ENABLE_SLAB", slab_keep.c_str(), 1);
  }
  // row-streaming kernel (output width >= 128, <= 64 channels per source)
  const ConvCase row_cases[] = {
      {"rows 3x3 64->64 @16x128 B2 +res dual",             2, 16, 128, 64, 1, 64, 64, 3, 1, true, true, true, false, false, false, 0},
      {"rows 3x3 32->32 @256x256 B1 +res dual (full size)", 1, 256, 256, 32, 1, 32, 32, 3, 1, true, true, true, false, false, false, 0},
      {"rows 1x1 32->32 @16x256 B2 elu only",              2, 16, 256, 32, 1, 32, 32, 1, 1, false, false, true, false, false, false, 0},
      {"rows 1x1 64->64 @8x128 B3 dual",                   3, 8, 128, 64, 1, 64, 64, 1, 1, false, true, true, false, false, false, 0},
      {"rows 3x3 64->128 D2S @16x128 B1 dual",             1, 16, 128, 64, 1, 128, 128, 3, 1, false, true, true, true, false, false, 0},
      {"rows 3x3 [32|32]->32 @32x256 B2 cat +res raw",     2, 32, 256, 32, 2, 32, 32, 3, 1, true, true, false, false, false, false, 0},
      {"rows 3x3 64->64 @128x128 B2 +res dual",            2, 128, 128, 64, 1, 64, 64, 3, 1, true, true, true, false, false, false, 0},
  };
  for (const ConvCase& c : row_cases) run_conv_case(c);
  // ELU-on-load (raw sources transformed in shared memory by the row kernel's extra warps)
  {
    const ConvCase e1 = {"rows elu-in 3x3 32->32 @256x256 B1 +res raw", 1, 256, 256, 32, 1, 32, 32, 3, 1, true, true, false, false, false, false, 0};
    const ConvCase e2 = {"rows elu-in 3x3 64->64 @128x128 B2 +res dual", 2, 128, 128, 64, 1, 64, 64, 3, 1, true, true, true, false, false, false, 0};
    const ConvCase e3 = {"rows elu-in 1x1 32->32 @16x256 B2 elu only", 2, 16, 256, 32, 1, 32, 32, 1, 1, false, false, true, false, false, false, 0};
    const ConvCase e4 = {"rows elu-in 1x1 64->64 @8x128 B3 elu only", 3, 8, 128, 64, 1, 64, 64, 1, 1, false, false, true, false, false, false, 0};
    const ConvCase e5 = {"rows elu-in(src0) 3x3 [32|32]->32 @32x256 B2 cat +res raw", 2, 32, 256, 32, 2, 32, 32, 3, 1, true, true, false, false, false, false, 0};
    g_elu_in = 1;
    run_conv_case(e1);
    run_conv_case(e2);
    run_conv_case(e3);
    run_conv_case(e4);
    run_conv_case(e5);
    setenv("VUB_ROWS_GRID", "5", 1);
    run_conv_case(e2);
    run_conv_case(e5);
    unsetenv("VUB_ROWS_GRID");
    g_elu_in = 3;
    run_conv_case(e5);
    g_elu_in = 0;
  }
  setenv("VUB_ROWS_GRID", "7", 1);  // few persistent CTAs: long row ranges that cross image boundaries, ring wraps
  {
    const ConvCase c1 = {"rows persistent grid=7: 3x3 32->32 @64x128 B6 +res dual", 6, 64, 128, 32, 1, 32, 32, 3, 1, true, true, true, false, false, false, 0};
    const ConvCase c2 = {"rows persistent grid=7: 1x1 32->32 @64x128 B6 dual", 6, 64, 128, 32, 1, 32, 32, 1, 1, false, true, true, false, false, false, 0};
    const ConvCase c3 = {"rows persistent grid=7: 3x3 64->64 @64x128 B5 +res dual", 5, 64, 128, 64, 1, 64, 64, 3, 1, true, true, true, false, false, false, 0};
    run_conv_case(c1);
    run_conv_case(c2);
    run_conv_case(c3);
  }
  unsetenv("VUB_ROWS_GRID");
  setenv("VUB_ROWS_ATM", "0", 1);  // A operand from shared memory (the default keeps it in tensor memory)
  for (const ConvCase& c : row_cases) run_conv_case(c);
  unsetenv("VUB_ROWS_ATM");
  setenv("VUB_DISABLE_ROWS", "1", 1);  // same wide shapes through the per-tap kernel
  run_conv_case(row_cases[0]);
  run_conv_case(row_cases[4]);
  unsetenv("VUB_DISABLE_ROWS");
  test_wn_prep();
  test_stem(2, 3, 32, 64);
  test_stem(3, 30, 16, 16);
  test_latent(3, 128, 4, 4);
  test_latent(2, 128, 8, 8);
  test_convert(2, 128, 4, 8);
  test_convert(3, 3, 16, 16);

  if (perf) {
    const int B = 16;
    for (int e : {0, 1}) {  // (raw, elu) pair I/O vs lean I/O: raw source with ELU on load, raw output only
      g_elu_in = e;
      perf_conv(e ? "lean 32->32 3x3 @256^2 +res raw, elu-in" : "pair 32->32 3x3 @256^2 +res dual", 64, 256, 256, 32, 1, 32, 3, 1, true, e == 0, false, 0);
      perf_conv(e ? "lean 64->64 3x3 @128^2 +res raw, elu-in" : "pair 64->64 3x3 @128^2 +res dual", 64, 128, 128, 64, 1, 64, 3, 1, true, e == 0, false, 0);
      perf_conv(e ? "lean [32|32]->32 3x3 @256^2 +res raw, elu-in" : "pair [32|32]->32 3x3 @256^2 +res dual", 64, 256, 256, 32, 2, 32, 3, 1, true, e == 0, false, 0);
      perf_conv(e ? "lean 32->32 1x1 @256^2 elu-in" : "pair 32->32 1x1 @256^2", 64, 256, 256, 32, 1, 32, 1, 1, false, false, false, 0);
    }
    g_elu_in = 0;
    for (int st : {0, -1}) {  // 0: library default (row-streaming where applicable); -1: per-tap kernel
      perf_conv("32->32 3x3 @256^2 +res dual", B, 256, 256, 32, 1, 32, 3, 1, true, true, false, st);
      perf_conv("[32|32]->32 3x3 @256^2 +res dual", B, 256, 256, 32, 2, 32, 3, 1, true, true, false, st);
      perf_conv("64->64 3x3 @128^2 +res dual", B, 128, 128, 64, 1, 64, 3, 1, true, true, false, st);
      perf_conv("128->128 3x3 @64^2 +res dual", B, 64, 64, 128, 1, 128, 3, 1, true, true, false, st);
      perf_conv("[128|128]->128 3x3 @64^2 +res dual", B, 64, 64, 128, 2, 128, 3, 1, true, true, false, st);
      perf_conv("128->256 3x3 D2S @64^2 dual", B, 64, 64, 128, 1, 256, 3, 1, false, true, true, st);
    }
    perf_conv("32->32 1x1 @256^2 elu", B, 256, 256, 32, 1, 32, 1, 1, false, false, false, 0);
    perf_conv("32->64 3x3 s2 @256^2 dual", B, 256, 256, 32, 1, 64, 3, 2, false, true, false, 0);
  }
  printf("%s: %d failed\n", g_fail ? "FAILED" : "ALL PASSED", g_fail);
  return g_fail;
}
