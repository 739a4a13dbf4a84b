// Torch-free GPU self-test of the kernels written after the round's GPU budget was nearly spent: the enable_alg_ext
// `init_scale` branches of the qdq kernels / fused grad-w epilogue and the histogram-select outlier loss (ar_outlier.cu).
// Starts in about a second (no Python, no torch import), so it fits in a very short GPU slot:
//
//     g++ -std=c++17 -O2 -ffp-contract=off tools/gpu_selftest.cpp -I/usr/local/cuda/include -Iauto_round_b200/csrc \
//         -Lauto_round_b200/csrc -lar_b200 -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread \
//         -Wl,-rpath,'$ORIGIN/../auto_round_b200/csrc' -o tools/gpu_selftest          (tools/build_selftest.sh)
//     tools/gpu_selftest            -> one PASS/FAIL line per check, exit code = number of failures
//
// Host-only source (every kernel lives in libar_b200.so), hence plain g++: without nvcc the `__device__` qualifiers of the
// math header expand to nothing and its structs run on the CPU.
//
// The expected values come from the PRODUCT's own math header compiled for the host in this same translation unit (the
// arrangement tests/test_host_math.py pins against the oracle on the CPU), so a PASS here closes the loop
// oracle == host math == device kernels.  Development tool, not part of the library.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#include "ar_qdq_math.cuh"

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e__ = (x);                                                                     \
    if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 99; } \
  } while (0)
#define AR(x)                                                                                 \
  do {                                                                                        \
    int rc__ = (x);                                                                           \
    if (rc__ != 0) { printf("FAIL %s -> %d: %s\n", #x, rc__, ar_last_error()); ++fails; }      \
  } while (0)

static int fails = 0;
static uint32_t rng_state = 12345u;
static float urand() { rng_state = rng_state * 1664525u + 1013904223u; return (rng_state >> 8) * (1.0f / 16777216.0f); }
static float nrand() { float s = 0; for (int i = 0; i < 12; ++i) s += urand(); return s - 6.f; }

static uint16_t h_bf16_bits(float f) { __nv_bfloat16 b = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &b, 2); return u; }
static float h_bf16_val(uint16_t u) { uint32_t w = ((uint32_t)u) << 16; float f; memcpy(&f, &w, 4); return f; }
static float h_bf16_round(float f) { return h_bf16_val(h_bf16_bits(f)); }

template <class T>
static T* dev(const std::vector<T>& h) { T* d = nullptr; cudaMalloc(&d, h.size() * sizeof(T)); cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); return d; }
template <class T>
static std::vector<T> host(const T* d, size_t n) { std::vector<T> h(n); cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost); return h; }

static void report(const char* name, bool ok, const char* detail = "") {
  printf("%s %s %s\n", ok ? "PASS" : "FAIL", name, detail);
  if (!ok) ++fails;
}

// host mirror of the kernels' group handling, same template as tests/host_math/math_host.cpp
template <class Ctx, bool FP4>
static void host_qdq(int bits, int g, long groups, const float* w, const float* v, const float* mn, const float* mx, const float* init,
                     float gscale, float thr, const float* gq, float* wq, float* scale, float* dv, float* dmin, float* dmax) {
  for (long gi_ = 0; gi_ < groups; ++gi_) {
    const float* wg = w + gi_ * g;
    ar::GroupIn gi;
    gi.thr = thr; gi.gscale = gscale; gi.plain = false;
    gi.mn = mn ? mn[gi_] : 1.f; gi.mx = mx ? mx[gi_] : 1.f;
    gi.has_init = (init != nullptr); gi.init = init ? init[gi_] : 1.f;
    if (FP4) { float m = 0.f; for (int i = 0; i < g; ++i) m = fmaxf(m, fabsf(wg[i])); gi.wmax = m; gi.wmin = 0.f; }
    else { float lo = 0.f, hi = 0.f; for (int i = 0; i < g; ++i) { lo = fminf(lo, wg[i]); hi = fmaxf(hi, wg[i]); } gi.wmin = lo; gi.wmax = hi; }
    Ctx ctx; ctx.init(bits); ctx.setup(gi);
    ar::GroupAcc acc;
    for (int i = 0; i < g; ++i) {
      const float vv = v[gi_ * g + i];
      wq[gi_ * g + i] = h_bf16_round(ctx.fwd(wg[i], vv));
      if (gq) { float d; ctx.bwd(wg[i], vv, gq[gi_ * g + i], d, acc); dv[gi_ * g + i] = d; }
    }
    scale[gi_] = ctx.scale_out();
    if (gq) { float a, b; ctx.finish(acc, gi, a, b); dmin[gi_] = a; dmax[gi_] = b; }
  }
}

static int test_init_scale(int dtype, int bits, int g, const char* name) {
  const int n = 64, k = 256;
  const long groups = (long)n * k / g;
  std::vector<float> w(n * k), v(n * k), mn(groups), mx(groups), init(groups), gq(n * k);
  std::vector<uint16_t> wb(n * k);
  for (int i = 0; i < n * k; ++i) { w[i] = h_bf16_round(0.05f * nrand()); wb[i] = h_bf16_bits(w[i]); v[i] = urand() - 0.5f; gq[i] = h_bf16_round(nrand()); }
  for (int i = 0; i < g; ++i) { w[i] = 0.f; wb[i] = 0; }                        // an all-zero group
  float amax = 0.f;
  for (float x : w) amax = fmaxf(amax, fabsf(x));
  const float gscale = (dtype == AR_DT_NV_FP4) ? 448.f * 6.f / amax : 0.f;
  for (long i = 0; i < groups; ++i) {
    mn[i] = 0.5f + 0.5f * urand(); mx[i] = 0.8f + 1.2f * urand();
    if (dtype == AR_DT_INT_SYM) {                                               // a plausible searched scale, bf16-valued, either sign
      float m = 0.f; for (int j = 0; j < g; ++j) m = fmaxf(m, fabsf(w[i * g + j]));
      init[i] = h_bf16_round(((i & 1) ? -1.f : 1.f) * (m / (float)(1 << (bits - 1))) * (0.8f + 0.4f * urand()));
      if (init[i] == 0.f) init[i] = 1.0013580322265625e-05f;
    } else {
      init[i] = (dtype == AR_DT_MX_FP4) ? ((i % 3 == 0) ? 0.5f : ((i % 3 == 1) ? 1.f : 2.f)) : 0.5f + 0.01f * (float)(i % 102);
    }
  }
  std::vector<float> e_wq(n * k), e_scale(groups), e_dv(n * k), e_dmin(groups, 0.f), e_dmax(groups);
  const float thr = 1e-5f;
  if (dtype == AR_DT_INT_SYM) host_qdq<ar::IntSym, false>(bits, g, groups, w.data(), v.data(), mn.data(), mx.data(), init.data(), gscale, thr, gq.data(), e_wq.data(), e_scale.data(), e_dv.data(), e_dmin.data(), e_dmax.data());
  else if (dtype == AR_DT_MX_FP4) host_qdq<ar::MxFp4, true>(bits, g, groups, w.data(), v.data(), nullptr, mx.data(), init.data(), gscale, thr, gq.data(), e_wq.data(), e_scale.data(), e_dv.data(), e_dmin.data(), e_dmax.data());
  else host_qdq<ar::NvFp4, true>(bits, g, groups, w.data(), v.data(), nullptr, mx.data(), init.data(), gscale, thr, gq.data(), e_wq.data(), e_scale.data(), e_dv.data(), e_dmin.data(), e_dmax.data());

  uint16_t* d_w = dev(wb);
  float *d_v = dev(v), *d_mn = dev(mn), *d_mx = dev(mx), *d_init = dev(init), *d_gq = dev(gq);
  std::vector<float> gsv(1, gscale);
  float* d_gs = dev(gsv);
  uint16_t *d_wq, *d_wmin, *d_wmax;
  float *d_dv, *d_dmin, *d_dmax;
  CK(cudaMalloc(&d_wq, n * k * 2)); CK(cudaMalloc(&d_wmin, groups * 2)); CK(cudaMalloc(&d_wmax, groups * 2));
  CK(cudaMalloc(&d_dv, n * k * 4)); CK(cudaMalloc(&d_dmin, groups * 4)); CK(cudaMalloc(&d_dmax, groups * 4));
  ar_qspec q{dtype, bits, g, n, k, thr, 2.0f, d_init};
  const bool is_int = (dtype == AR_DT_INT_SYM);
  if (is_int) AR(ar_group_minmax(&q, d_w, d_wmin, d_wmax, nullptr));
  AR(ar_qdq_fwd(&q, d_w, d_v, is_int ? d_mn : nullptr, d_mx, is_int ? d_wmin : nullptr, is_int ? d_wmax : nullptr,
                dtype == AR_DT_NV_FP4 ? d_gs : nullptr, d_wq, nullptr, nullptr, nullptr));
  AR(ar_qdq_bwd(&q, d_w, d_v, is_int ? d_mn : nullptr, d_mx, is_int ? d_wmin : nullptr, is_int ? d_wmax : nullptr,
                dtype == AR_DT_NV_FP4 ? d_gs : nullptr, d_gq, d_dv, is_int ? d_dmin : nullptr, d_dmax, 0, nullptr));
  CK(cudaDeviceSynchronize());
  auto g_wq = host(d_wq, (size_t)n * k);
  auto g_dv = host(d_dv, (size_t)n * k);
  auto g_dmax = host(d_dmax, (size_t)groups);
  long bad_wq = 0, bad_dv = 0, bad_dm = 0;
  float dm_ref = 0.f;
  for (float x : e_dmax) if (std::isfinite(x)) dm_ref = fmaxf(dm_ref, fabsf(x));
  for (int i = 0; i < n * k; ++i) {
    if (h_bf16_val(g_wq[i]) != e_wq[i]) ++bad_wq;
    if (g_dv[i] != e_dv[i]) ++bad_dv;
  }
  for (long i = 0; i < groups; ++i) if (std::isfinite(e_dmax[i]) && fabsf(g_dmax[i] - e_dmax[i]) > 2e-5f * dm_ref) ++bad_dm;
  char buf[160];
  snprintf(buf, sizeof buf, "(wq mismatches %ld, dv mismatches %ld, dmax out of tol %ld of %ld groups)", bad_wq, bad_dv, bad_dm, groups);
  char nm[96];
  snprintf(nm, sizeof nm, "init_scale qdq fwd/bwd %s", name);
  report(nm, bad_wq == 0 && bad_dv == 0 && bad_dm == 0, buf);
  if (is_int) {
    auto g_dmin = host(d_dmin, (size_t)groups);
    long nz = 0;
    for (float x : g_dmin) if (x != 0.f) ++nz;
    snprintf(nm, sizeof nm, "init_scale d(min_scale) == 0 %s", name);
    report(nm, nz == 0);
  }

  // fused grad-w epilogue with an init scale: dV from the tcgen05 GEMM epilogue == host backward of Gq = dY^T X
  const int T = 256;
  std::vector<uint16_t> dyb((size_t)T * n), xb((size_t)T * k);
  std::vector<float> dy((size_t)T * n), x((size_t)T * k), gq2((size_t)n * k, 0.f);
  for (size_t i = 0; i < dy.size(); ++i) { dy[i] = h_bf16_round(0.1f * nrand()); dyb[i] = h_bf16_bits(dy[i]); }
  for (size_t i = 0; i < x.size(); ++i) { x[i] = h_bf16_round(nrand()); xb[i] = h_bf16_bits(x[i]); }
  for (int t = 0; t < T; ++t)
    for (int r = 0; r < n; ++r) {
      const float a = dy[(size_t)t * n + r];
      for (int c = 0; c < k; ++c) gq2[(size_t)r * k + c] += a * x[(size_t)t * k + c];
    }
  std::vector<float> e_dv2(n * k), e_dmin2(groups, 0.f), e_dmax2(groups), tmpq(n * k), tmps(groups);
  if (dtype == AR_DT_INT_SYM) host_qdq<ar::IntSym, false>(bits, g, groups, w.data(), v.data(), mn.data(), mx.data(), init.data(), gscale, thr, gq2.data(), tmpq.data(), tmps.data(), e_dv2.data(), e_dmin2.data(), e_dmax2.data());
  else if (dtype == AR_DT_MX_FP4) host_qdq<ar::MxFp4, true>(bits, g, groups, w.data(), v.data(), nullptr, mx.data(), init.data(), gscale, thr, gq2.data(), tmpq.data(), tmps.data(), e_dv2.data(), e_dmin2.data(), e_dmax2.data());
  else host_qdq<ar::NvFp4, true>(bits, g, groups, w.data(), v.data(), nullptr, mx.data(), init.data(), gscale, thr, gq2.data(), tmpq.data(), tmps.data(), e_dv2.data(), e_dmin2.data(), e_dmax2.data());
  uint16_t *d_dy = dev(dyb), *d_x = dev(xb);
  AR(ar_fq_linear_bwd_dw(&q, d_dy, d_x, T, d_w, d_v, is_int ? d_mn : nullptr, d_mx, is_int ? d_wmin : nullptr, is_int ? d_wmax : nullptr,
                         dtype == AR_DT_NV_FP4 ? d_gs : nullptr, d_dv, 0, is_int ? d_dmin : nullptr, d_dmax, 0, nullptr));
  CK(cudaDeviceSynchronize());
  auto g_dv2 = host(d_dv, (size_t)n * k);
  float vmax = 0.f, verr = 0.f;
  for (int i = 0; i < n * k; ++i) { vmax = fmaxf(vmax, fabsf(e_dv2[i])); verr = fmaxf(verr, fabsf(g_dv2[i] - e_dv2[i])); }
  snprintf(buf, sizeof buf, "(max |dV err| %.3g of max |dV| %.3g)", verr, vmax);
  snprintf(nm, sizeof nm, "init_scale fused grad-w epilogue %s", name);
  report(nm, verr <= 2e-3f * vmax, buf);
  return 0;
}

static int test_outlier(int rows, int cols, bool masked) {
  const long numel = (long)rows * cols;
  std::vector<uint16_t> pb(numel), rb(numel);
  std::vector<float> p(numel), r(numel);
  std::vector<uint8_t> mask(rows, 1);
  for (long i = 0; i < numel; ++i) {
    r[i] = h_bf16_round(nrand()); p[i] = h_bf16_round(r[i] + 0.05f * nrand());
    rb[i] = h_bf16_bits(r[i]); pb[i] = h_bf16_bits(p[i]);
  }
  p[5] = h_bf16_round(p[5] + 2.f); pb[5] = h_bf16_bits(p[5]);
  if (masked) for (int i = 0; i < rows; ++i) mask[i] = (urand() > 0.1f) ? 1 : 0;
  const long k = std::max(1L, numel / 1000);
  std::vector<uint32_t> bits(numel);
  for (long i = 0; i < numel; ++i) bits[i] = (uint32_t)h_bf16_bits(p[i] - r[i]) & 0x7fffu;
  std::vector<uint32_t> sorted(bits);
  std::nth_element(sorted.begin(), sorted.begin() + (k - 1), sorted.end(), std::greater<uint32_t>());
  const uint32_t thr = sorted[k - 1];
  long above = 0, at = 0;
  for (uint32_t b : bits) { above += (b > thr); at += (b == thr); }
  const long need = k - above;
  // loss bounds over every admissible choice of `need` tie members (they share the bf16 |diff| but not the fp32 one)
  double base = 0.0;
  std::vector<double> tie_sq;
  for (long i = 0; i < numel; ++i) {
    const bool on = mask[i / cols] != 0;
    const double d = (double)p[i] - (double)r[i];
    if (bits[i] < thr) base += on ? d * d : 0.0;
    else if (bits[i] == thr) tie_sq.push_back(on ? d * d : 0.0);
  }
  std::sort(tie_sq.begin(), tie_sq.end());
  double lo = base, hi = base;
  for (long i = 0; i < (long)tie_sq.size() - need; ++i) lo += tie_sq[i];
  for (long i = need; i < (long)tie_sq.size(); ++i) hi += tie_sq[i];

  uint16_t *d_p = dev(pb), *d_r = dev(rb), *d_dp;
  uint8_t* d_m = dev(mask);
  uint32_t *d_hist, *d_sel;
  double* d_loss;
  CK(cudaMalloc(&d_dp, numel * 2)); CK(cudaMalloc(&d_hist, 32768 * 4)); CK(cudaMalloc(&d_sel, 16)); CK(cudaMalloc(&d_loss, 8));
  CK(cudaMemset(d_hist, 0, 32768 * 4)); CK(cudaMemset(d_sel, 0, 16)); CK(cudaMemset(d_loss, 0, 8));
  double got_loss[2];
  for (int rep = 0; rep < 2; ++rep) {                      // twice: the histogram and the tie counter must re-arm themselves
    CK(cudaMemset(d_loss, 0, 8));
    AR(ar_absdiff_hist(d_p, d_r, numel, d_hist, nullptr));
    AR(ar_topk_threshold(d_hist, k, d_sel, nullptr));
    AR(ar_mse_outlier_fwd_bwd(d_p, d_r, masked ? d_m : nullptr, rows, cols, 1000.f, d_sel, d_loss, d_dp, nullptr));
    CK(cudaDeviceSynchronize());
    got_loss[rep] = host(d_loss, 1)[0];
  }
  auto sel = host(d_sel, 3);
  auto hist = host(d_hist, 32768);
  auto dp = host(d_dp, (size_t)numel);
  long hist_nz = 0;
  for (uint32_t c : hist) hist_nz += (c != 0);
  char buf[200], nm[96];
  snprintf(nm, sizeof nm, "outlier select %dx%d%s", rows, cols, masked ? " masked" : "");
  snprintf(buf, sizeof buf, "(thr %u/%u need %u/%ld ties seen %u of %ld, hist nonzero %ld)", sel[0], thr, sel[1], need, sel[2], at, hist_nz);
  report(nm, sel[0] == thr && (long)sel[1] == need && (long)sel[2] == at && hist_nz == 0, buf);
  snprintf(nm, sizeof nm, "outlier loss %dx%d%s", rows, cols, masked ? " masked" : "");
  bool ok = true;
  for (int rep = 0; rep < 2; ++rep) ok = ok && got_loss[rep] >= lo * (1 - 1e-5) && got_loss[rep] <= hi * (1 + 1e-5);
  snprintf(buf, sizeof buf, "(sum %.9g / %.9g in [%.9g, %.9g])", got_loss[0], got_loss[1], lo, hi);
  report(nm, ok, buf);
  // gradient on every element that is not a tie member: ((1000/numel) * (2 x)) * sign(d), x = |d| m keep
  const float up = 1000.f / (float)numel;
  long bad = 0, dropped_above = 0;
  for (long i = 0; i < numel; ++i) {
    if (bits[i] == thr) continue;
    const bool keep = bits[i] < thr, on = mask[i / cols] != 0;
    const float d = p[i] - r[i];
    const float xx = (on && keep) ? fabsf(d) : 0.f;
    const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    const float e = h_bf16_round((up * (2.f * xx)) * sg);
    if (h_bf16_val(dp[i]) != e) ++bad;
    if (!keep) ++dropped_above;
  }
  snprintf(nm, sizeof nm, "outlier grad %dx%d%s", rows, cols, masked ? " masked" : "");
  snprintf(buf, sizeof buf, "(%ld mismatches; %ld elements above the threshold)", bad, dropped_above);
  report(nm, bad == 0 && dropped_above == above, buf);
  return 0;
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no CUDA device\n"); return 98; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("device: %s (sm_%d%d), lib version %d\n", prop.name, prop.major, prop.minor, ar_version());
  if (test_outlier(64, 64, true)) return 99;
  if (test_outlier(1000, 128, false)) return 99;
  if (test_outlier(4096, 4096, true)) return 99;
  if (test_init_scale(AR_DT_INT_SYM, 2, 32, "int_sym w2 g32")) return 99;
  if (test_init_scale(AR_DT_INT_SYM, 4, 128, "int_sym w4 g128")) return 99;
  if (test_init_scale(AR_DT_MX_FP4, 4, 32, "mx_fp4")) return 99;
  if (test_init_scale(AR_DT_NV_FP4, 4, 16, "nv_fp4")) return 99;
  printf("%s: %d failure(s)\n", fails ? "SELFTEST FAILED" : "SELFTEST OK", fails);
  return fails;
}
