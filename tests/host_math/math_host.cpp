// TEST INFRASTRUCTURE: the product's device math header (auto_round_b200/csrc/ar_qdq_math.cuh) compiled as plain host C++
// (g++, no nvcc, no GPU) behind a C entry point, so that the fake-quant forward/backward formulas -- including the
// enable_alg_ext init_scale branches that have not run on hardware yet -- can be checked against the oracle in the CPU
// test tier.  The CUDA headers provide host versions of the fp16 / bf16 / e4m3 conversions; the only shims are
// __uint_as_float, __float_as_uint and __fmaf_rn.  Group handling mirrors qdq_fwd_kernel / qdq_bwd_kernel (ar_qdq.cu): wmin/wmax clamped at 0, fp4 amax,
// sequential (not shuffle-tree) group sums.
#include <cmath>
#include <cstdint>
#include <cstring>
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#include "../../auto_round_b200/csrc/ar_qdq_math.cuh"

namespace {
template <class Ctx, bool FP4>
void run(int bits, int g, long groups, const float* w, const float* v, const float* mn, const float* mx, const float* init,
         float gscale, float thr, const float* gq, float* wq, float* scale, float* zp, float* dv, float* dmin, float* dmax) {
  for (long gi_ = 0; gi_ < groups; ++gi_) {
    const float* wg = w + gi_ * g;
    ar::GroupIn gi;
    gi.thr = thr; gi.gscale = gscale; gi.plain = (mn == nullptr) && (mx == nullptr);
    gi.mn = mn ? mn[gi_] : 1.f; gi.mx = mx ? mx[gi_] : 1.f;
    gi.has_init = (init != nullptr); gi.init = init ? init[gi_] : 1.f;
    if (FP4) {
      float m = 0.f;
      for (int i = 0; i < g; ++i) m = fmaxf(m, fabsf(wg[i]));
      gi.wmax = m; gi.wmin = 0.f;
    } else {
      float lo = 0.f, hi = 0.f;
      for (int i = 0; i < g; ++i) { lo = fminf(lo, wg[i]); hi = fmaxf(hi, wg[i]); }
      gi.wmin = lo; gi.wmax = hi;
    }
    Ctx ctx; ctx.init(bits); ctx.setup(gi);
    ar::GroupAcc acc;
    for (int i = 0; i < g; ++i) {
      const float vv = v ? v[gi_ * g + i] : 0.f;
      if (wq) wq[gi_ * g + i] = ar::bf16_round(ctx.fwd(wg[i], vv));
      if (gq) { float d; ctx.bwd(wg[i], vv, gq[gi_ * g + i], d, acc); dv[gi_ * g + i] = d; }
    }
    if (scale) scale[gi_] = ctx.scale_out();
    if (zp) zp[gi_] = ctx.zp_out();
    if (gq) { float a, b; ctx.finish(acc, gi, a, b); if (dmin) dmin[gi_] = a; if (dmax) dmax[gi_] = b; }
  }
}
}  // namespace

// dtype: 0 int_sym, 1 int_asym, 2 mx_fp4, 3 nv_fp4 (AR_DT_*).  w holds bf16 values as fp32.  Any output may be NULL.
extern "C" int host_qdq(int dtype, int bits, int g, long groups, const float* w, const float* v, const float* mn, const float* mx,
                        const float* init, float gscale, float thr, const float* gq, float* wq, float* scale, float* zp,
                        float* dv, float* dmin, float* dmax) {
  switch (dtype) {
    case 0: run<ar::IntSym, false>(bits, g, groups, w, v, mn, mx, init, gscale, thr, gq, wq, scale, zp, dv, dmin, dmax); return 0;
    case 1: run<ar::IntAsym, false>(bits, g, groups, w, v, mn, mx, init, gscale, thr, gq, wq, scale, zp, dv, dmin, dmax); return 0;
    case 2: run<ar::MxFp4, true>(bits, g, groups, w, v, mn, mx, init, gscale, thr, gq, wq, scale, zp, dv, dmin, dmax); return 0;
    case 3: run<ar::NvFp4, true>(bits, g, groups, w, v, mn, mx, init, gscale, thr, gq, wq, scale, zp, dv, dmin, dmax); return 0;
  }
  return -1;
}

// Exhaustive check of ar::div_exact (the division-free w / s of the hot kernels): every finite bf16 numerator against every
// finite fp16-valued scale with |s| >= 1e-5 (the q_scale_thresh clip guarantees that bound) against the IEEE quotient.
// Returns the number of pairs that are neither bit-identical nor "both below 2^-80 in magnitude" (tiny / zero numerators:
// invisible in round(w/s + V), see ar_qdq_math.cuh); *pairs = pairs checked, *exact = bit-identical pairs,
// *max_inexact_w = the largest |w| of a pair that is not bit-identical (must be far below any real weight).
extern "C" long host_div_exact_check(long* pairs, long* exact, float* max_inexact_w) {
  long bad = 0, tot = 0, same = 0;
  float worst = 0.f;
#pragma omp parallel for reduction(+ : bad, tot, same) reduction(max : worst) schedule(dynamic, 64)
  for (int hs = 0; hs < 65536; ++hs) {
    const int e = (hs >> 10) & 31, m = hs & 1023;
    if (e == 31) continue;
    float s = (e == 0) ? ldexpf((float)m, -24) : ldexpf((float)(m | 1024), e - 25);
    if (hs & 0x8000) s = -s;
    if (!(fabsf(s) >= 1.0e-5f)) continue;
    volatile float one = 1.f;
    const float rs = one / s;
    for (int wb = 0; wb < 65536; ++wb) {
      const float w = __uint_as_float(((uint32_t)wb) << 16);
      if (std::isnan(w) || std::isinf(w)) continue;
      if (fabsf(w) > 65504.f * 128.f) continue;            // an fp16 scale cannot represent such a group (s = max|w| / maxq)
      volatile float wv = w, sv = s;
      const float ref = wv / sv;
      const float got = ar::div_exact(w, s, rs);
      uint32_t a, b;
      memcpy(&a, &got, 4);
      memcpy(&b, &ref, 4);
      ++tot;
      if (a == b) { ++same; continue; }
      const float tiny = ldexpf(1.f, -80);
      if (fabsf(got) < tiny && fabsf(ref) < tiny) { worst = fmaxf(worst, fabsf(w)); continue; }
      ++bad;
    }
  }
  if (pairs) *pairs = tot;
  if (exact) *exact = same;
  if (max_inexact_w) *max_inexact_w = worst;
  return bad;
}
