// Device-side fake-quant numerics shared by the standalone qdq kernels (ar_qdq.cu) and the fused
// epilogue of the dW GEMM (ar_gemm.cu).  One struct per AutoRound data type:
//
//     Ctx::setup(group inputs)        once per quantisation group
//     ctx.fwd(w, v)                   -> dequantised value (fp32, before the bf16 cast)
//     ctx.bwd(w, v, gq, dv, acc)      -> dL/dV for this element, group partial sums in `acc`
//     ctx.finish(acc, dmin, dmax)     -> dL/d(min_scale), dL/d(max_scale) of the group
//
// Every expression mirrors the reference's torch op sequence in fp32 (IEEE div, half-to-even rint, no FMA
// contraction: the library is compiled with -fmad=false) so that round(W/s + V) is bit-identical:
//   int_sym  auto_round/data_type/int.py:165-238      int_asym auto_round/data_type/int.py:241-298
//   mx_fp4   auto_round/data_type/mxfp.py:49-85,233-291   nv_fp4 auto_round/data_type/nvfp.py:26-98
// The backward formulas are the closed form of torch autograd over those graphs (straight-through
// estimators: auto_round/data_type/utils.py:314-365); derivation in DESIGN.md section 4.
#pragma once
#include "ar_common.cuh"

namespace ar {

struct GroupIn {
  float wmin, wmax;  // int types: min/max of the group clamped at 0 (bf16 values); fp4: wmax = amax|W|
  float mn, mx;      // min_scale / max_scale of the group (1 when absent)
  float gscale;      // NVFP4 per-tensor global scale
  float thr;         // q_scale_thresh
  bool plain = false;  // no min/max_scale tensors (RTN, iters == 0): the reference's range math then stays in the
                       // weight dtype (bf16) because 0-dim / python scales do not promote (int.py:278-286)
  float init = 1.f;    // enable_alg_ext: searched initial scale of the group (ar_qspec::init_scale)
  bool has_init = false;
};

struct GroupAcc {
  float a = 0.f, b = 0.f;  // meaning is per data type (see bwd/finish)
};

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// w / s without the IEEE division sequence on the hot path: with rs = RN(1/s) (one real division per GROUP),
//   q0 = RN(w*rs);  e = w - s*q0 (exact, FMA);  q = RN(q0 + e*rs)
// is the correctly rounded quotient (Markstein) whenever the residual e does not underflow.  For the operands that occur
// here -- w a bf16 weight, s an fp16-valued scale with |s| >= 1e-5 -- tests/test_div_exact.py checks ALL pairs against the
// IEEE quotient: bit-identical for every |w| >= 2^-100 (any real weight); below that (and for w = +-0, where only the
// sign of the zero can differ) both quotients are smaller than 2^-80 in magnitude, i.e. invisible in round(w/s + V): V moves
// on a grid of >= 1e-3 steps around +0, so q + V == V (or q rounds to 0 when V == 0) for either value.  No branch, no guard:
// an fp16-valued scale also bounds |w| <= 65504 * maxq, so nothing overflows.
__device__ __forceinline__ float div_exact(float w, float s, float rs) {
  const float q0 = w * rs;
  const float e = __fmaf_rn(-q0, s, w);
  return __fmaf_rn(e, rs, q0);
}

// ------------------------------------------------------------------------------------------------ int_sym
struct IntSym {
  float kMaxq;    // 2^(bits-1)
  float kInvMaxq; // 2^-(bits-1): x / kMaxq == x * kInvMaxq exactly (power of two)
  float s;        // scale: fp16-rounded, threshold-clipped, as fp32
  float rs;       // RN(1/s), see div_exact
  float route_mx; // d s_raw / d max_scale
  float route_mn; // d s_raw / d min_scale

  __device__ __forceinline__ void init(int bits) { kMaxq = (float)(1 << (bits - 1)); kInvMaxq = 1.f / kMaxq; }
  __device__ __forceinline__ void setup(const GroupIn& g) {
    if (g.has_init) {                                  // int.py:201-216: scale = fp16(init_scale * max_scale), then the clip
      const float s_raw = f16_round(g.init * g.mx);
      const float thr = f16_round(g.thr);
      bool pass;
      if (s_raw < 0.f) { s = fminf(s_raw, -thr); pass = (s_raw <= -thr); }
      else             { s = fmaxf(s_raw, thr);  pass = (s_raw >= thr); }
      route_mx = pass ? g.init : 0.f;
      route_mn = 0.f;                                  // min_scale does not enter this graph
      rs = 1.f / s;
      return;
    }
    const float lo = -(g.wmin * g.mn);
    const float hi = g.wmax * g.mx;
    const float sgn = (hi < lo) ? 1.f : -1.f;        // "full range": +max maps to -maxq  (int.py:228-230)
    const float s_raw = f16_round((sgn * fmaxf(hi, lo)) * kInvMaxq);
    const float thr = f16_round(g.thr);               // clamp runs in the fp16 tensor's dtype
    bool pass;
    if (s_raw < 0.f) { s = fminf(s_raw, -thr); pass = (s_raw <= -thr); }
    else             { s = fmaxf(s_raw, thr);  pass = (s_raw >= thr); }
    // torch.max(hi, lo) backward: winner takes all, ties split evenly
    const float whi = hi > lo ? 1.f : (hi == lo ? 0.5f : 0.f);
    const float base = pass ? (sgn * kInvMaxq) : 0.f;
    route_mx = base * whi * g.wmax;
    route_mn = base * (1.f - whi) * (-g.wmin);
    rs = 1.f / s;
  }
  __device__ __forceinline__ float code(float w, float v) const {   // q in [-maxq, maxq-1]
    return clampf(rintf(div_exact(w, s, rs) + v), -kMaxq, kMaxq - 1.f);
  }
  __device__ __forceinline__ float fwd(float w, float v) const { return s * code(w, v); }
  // 4-bit wire code of the fake-quant value (data-parallel exchange of the next Wq): fwd == a * (enc - off), bits <= 4
  __device__ __forceinline__ uint32_t enc(float w, float v) const { return (uint32_t)(int)(code(w, v) + kMaxq); }
  __device__ __forceinline__ void dec_params(float& a, float& off) const { a = s; off = kMaxq; }
  __device__ __forceinline__ void bwd(float w, float v, float gq, float& dv, GroupAcc& acc) const {
    const float ws = div_exact(w, s, rs);
    const float r = rintf(ws + v);
    const float q = clampf(r, -kMaxq, kMaxq - 1.f);
    const float dt = (q == r) ? gq * s : 0.f;                        // clamp not active  <=>  q == r
    dv = dt;
    acc.a += gq * q - dt * (ws * rs);                                // d L / d s (a group sum whose SIGN is used: ws/s ~ ws*rs)
  }
  __device__ __forceinline__ void finish(const GroupAcc& acc, const GroupIn&, float& dmin, float& dmax) const {
    dmax = acc.a * route_mx;
    dmin = acc.a * route_mn;
  }
  __device__ __forceinline__ float scale_out() const { return s; }
  __device__ __forceinline__ float zp_out() const { return kMaxq; }
};

// ----------------------------------------------------------------------------------------------- int_asym
struct IntAsym {
  float kMaxq;    // 2^bits - 1
  float s, rs, zp, lo;
  bool pass;

  __device__ __forceinline__ void init(int bits) { kMaxq = (float)((1 << bits) - 1); }
  __device__ __forceinline__ void setup(const GroupIn& g) {
    lo = g.wmin * g.mn;
    const float hi = g.wmax * g.mx;
    const float s_raw = g.plain ? f16_round(bf16_round(bf16_round(hi - lo) / kMaxq)) : f16_round((hi - lo) / kMaxq);
    const float thr = f16_round(g.thr);
    s = fmaxf(s_raw, thr);
    pass = (s_raw >= thr);
    rs = 1.f / s;
    zp = rintf((-lo) / s);
  }
  __device__ __forceinline__ float code(float w, float v) const {   // q in [0, maxq]
    return clampf(rintf(div_exact(w, s, rs) + v) + zp, 0.f, kMaxq);
  }
  __device__ __forceinline__ float fwd(float w, float v) const { return s * (code(w, v) - zp); }
  __device__ __forceinline__ uint32_t enc(float w, float v) const { return (uint32_t)(int)code(w, v); }
  __device__ __forceinline__ void dec_params(float& a, float& off) const { a = s; off = zp; }
  __device__ __forceinline__ void bwd(float w, float v, float gq, float& dv, GroupAcc& acc) const {
    const float ws = div_exact(w, s, rs);
    const float u = rintf(ws + v) + zp;
    const float q = clampf(u, 0.f, kMaxq);
    const float gs = gq * s;
    const float dt = (q == u) ? gs : 0.f;                            // clamp not active  <=>  q == u
    dv = dt;
    acc.a += gq * (q - zp) - dt * (ws * rs);                         // d L / d s  (direct + through W/s; sign only)
    acc.b += dt - gs;                                                // d L / d zp (clamped elements only)
  }
  __device__ __forceinline__ void finish(const GroupAcc& acc, const GroupIn& g, float& dmin, float& dmax) const {
    const float dzp = acc.b;
    // zp = round_ste(-lo / s):  d(-lo) = dzp / s ,  d s += -dzp * ((-lo / s) / s)
    const float ds = acc.a - dzp * (((-lo) / s) / s);
    const float ds_raw = pass ? ds : 0.f;
    const float dhi = ds_raw / kMaxq;
    const float dlo = -(dzp / s) - ds_raw / kMaxq;
    dmax = dhi * g.wmax;
    dmin = dlo * g.wmin;
  }
  __device__ __forceinline__ float scale_out() const { return s; }
  __device__ __forceinline__ float zp_out() const { return zp; }
};

// ------------------------------------------------------------------------------------------------- mx_fp4
__device__ __forceinline__ float py_mod2(float x) {   // torch.remainder(x, 2): result takes the divisor's sign
  float r = fmodf(x, 2.f);
  if (r != 0.f && r < 0.f) r += 2.f;
  return r;
}

// E2M1 element rounding, auto_round/data_type/mxfp.py:49-85 with ebits=2, mbits=3, max_norm=6
__device__ __forceinline__ float mx_quant_element(float t) {
  const float a0 = fabsf(t) + (t == 0.f ? 1.f : 0.f);
  const float p = a0 >= 4.f ? 4.f : (a0 >= 2.f ? 2.f : 1.f);          // 2^max(floor(log2 a0), 0) for a0 <= 6
  const float y = t / p * 2.f;
  const float a = fabsf(y);
  const float tie = (py_mod2(a - 0.5f) == 0.f) ? 1.f : 0.f;
  const float sg = (y > 0.f) ? 1.f : (y < 0.f ? -1.f : 0.f);
  const float r = sg * (floorf(a + 0.5f) - tie);
  return clampf(r / 2.f * p, -6.f, 6.f);
}

// E2M1 value (0, +-0.5 .. +-6, signed zero kept) <-> 4-bit wire code: sign << 3 | index into {0, .5, 1, 1.5, 2, 3, 4, 6}
__device__ __forceinline__ uint32_t e2m1_enc(float val) {
  const float a2 = fabsf(val) * 2.f;                       // 0 1 2 3 4 6 8 12
  const uint32_t idx = a2 < 4.5f ? (uint32_t)(int)a2 : (a2 < 7.f ? 5u : (a2 < 10.f ? 6u : 7u));
  return idx | ((__float_as_uint(val) >> 31) << 3);
}
__device__ __forceinline__ float e2m1_dec(uint32_t nib) {
  const uint32_t i = nib & 7u;
  const float mag = i < 5u ? 0.5f * (float)i : (i == 5u ? 3.f : (i == 6u ? 4.f : 6.f));
  return (nib & 8u) ? -mag : mag;
}

struct MxFp4 {
  __device__ __forceinline__ void init(int) {}
  float s;      // 2^e
  float rs;     // 2^-e: w / s == w * rs exactly (power of two)
  float e;      // shared exponent (after -emax and clamp)
  float m;      // amax * max_scale
  bool pass;

  __device__ __forceinline__ void setup(const GroupIn& g) {
    m = g.wmax * (g.has_init ? g.init * g.mx : g.mx);   // max_val *= init_scale * max_scale (mxfp.py:262-266)
    const float e_raw = (m == 0.f) ? 1.f : log2f(m);
    const float ef = floorf(e_raw) - 2.f;
    e = clampf(ef, -127.f, 127.f);
    pass = (ef >= -127.f) && (ef <= 127.f) && (m != 0.f);
    s = ldexpf(1.f, (int)e);
    rs = ldexpf(1.f, -(int)e);
  }
  __device__ __forceinline__ float fwd(float w, float v) const {
    const float t = clampf(w * rs + v, -6.f, 6.f);
    return mx_quant_element(t) * s;
  }
  __device__ __forceinline__ uint32_t enc(float w, float v) const { return e2m1_enc(mx_quant_element(clampf(w * rs + v, -6.f, 6.f))); }
  __device__ __forceinline__ void dec_params(float& a, float& off) const { a = s; off = 0.f; }
  __device__ __forceinline__ void bwd(float w, float v, float gq, float& dv, GroupAcc& acc) const {
    const float ws = w * rs;
    const float t = ws + v;
    const bool in = (t >= -6.f) && (t <= 6.f);
    const float tc = clampf(t, -6.f, 6.f);
    const float o = mx_quant_element(tc);
    // d o / d tc through quant_element's graph: 1 below |tc| < 1 (0 at exactly 0), o/tc above
    const float f = (fabsf(tc) >= 1.f) ? (o / tc) : (tc != 0.f ? 1.f : 0.f);
    const float dt = in ? gq * s * f : 0.f;
    dv = dt;
    acc.a += gq * o - dt * (ws * rs);                                 // d L / d s
  }
  __device__ __forceinline__ void finish(const GroupAcc& acc, const GroupIn& g, float& dmin, float& dmax) const {
    // s = 2^e, e = floor_ste(log2 m) - 2  =>  ds/dm = s / m ;  m = amax * max_scale
    dmax = pass ? acc.a * (s / m) * g.wmax : 0.f;
    if (g.has_init) dmax *= g.init;
    dmin = 0.f;
  }
  __device__ __forceinline__ float scale_out() const { return e; }
  __device__ __forceinline__ float zp_out() const { return 0.f; }
};

// ------------------------------------------------------------------------------------------------- nv_fp4
// cast_to_fp4, auto_round/data_type/nvfp.py:26-39
__device__ __forceinline__ float nv_cast_to_fp4(float x) {
  const float sg = (x > 0.f) ? 1.f : (x < 0.f ? -1.f : 0.f);
  const float a = fabsf(x);
  float r;
  if (a < 2.f) r = rintf(2.f * a) / 2.f;
  else if (a < 4.f) r = rintf(a);
  else r = 2.f * rintf(a / 2.f);
  return clampf(r, -6.f, 6.f) * sg;
}

struct NvFp4 {
  __device__ __forceinline__ void init(int) {}
  float sc;     // e4m3-valued block scale (fp32)
  float inv;    // 1 / (sc / gscale)   (0 when sc == 0)
  float rinv;   // 1 / inv
  float route;  // d inv / d max_scale

  __device__ __forceinline__ void setup(const GroupIn& g) {
    const float vmax = g.wmax * (g.has_init ? g.mx * g.init : g.mx);   // scale_coeff = max_scale * init_scale (nvfp.py:93-97)
    const float sc_raw = g.gscale * (vmax * 0.16666667163372040f);     // get_reciprocal(6.0) as fp32
    const float sc_c = clampf(sc_raw, -448.f, 448.f);
    sc = e4m3_bits_to_f32(f32_to_e4m3_bits(sc_c));
    const float rg = (g.gscale == 0.f) ? 0.f : 1.f / g.gscale;
    const float prod = sc * rg;
    inv = (prod == 0.f) ? 0.f : 1.f / prod;
    rinv = (inv == 0.f) ? 0.f : 1.f / inv;
    const bool pass = (sc_raw >= -448.f) && (sc_raw <= 448.f);
    // inv = 1/prod, prod = sc*rg, sc = gs*vmax/6 (STE through e4m3), vmax = amax*mx
    route = (prod != 0.f && pass) ? (-(inv / prod)) * rg * g.gscale * 0.16666667163372040f * g.wmax : 0.f;
    if (g.has_init) route *= g.init;
  }
  __device__ __forceinline__ float fwd(float w, float v) const {
    const float x = clampf(w * inv + v, -6.f, 6.f);
    return nv_cast_to_fp4(x) * rinv;
  }
  __device__ __forceinline__ uint32_t enc(float w, float v) const { return e2m1_enc(nv_cast_to_fp4(clampf(w * inv + v, -6.f, 6.f))); }
  __device__ __forceinline__ void dec_params(float& a, float& off) const { a = rinv; off = 0.f; }
  __device__ __forceinline__ void bwd(float w, float v, float gq, float& dv, GroupAcc& acc) const {
    const float x = w * inv + v;
    const bool in = (x >= -6.f) && (x <= 6.f);
    const float xc = clampf(x, -6.f, 6.f);
    const float c = nv_cast_to_fp4(xc);
    const float dx = (in && xc != 0.f) ? gq * rinv : 0.f;            // d cast/d x = sign(x)^2
    dv = dx;
    // d L / d inv = dx*w (through x)  -  gq*c * rinv/inv (through 1/inv)
    acc.a += dx * w - ((inv != 0.f) ? gq * c * (rinv / inv) : 0.f);
  }
  __device__ __forceinline__ void finish(const GroupAcc& acc, const GroupIn&, float& dmin, float& dmax) const {
    dmax = acc.a * route;   // reference yields NaN (0*inf) for an all-zero group; we define it as 0
    dmin = 0.f;
  }
  __device__ __forceinline__ float scale_out() const { return sc; }
  __device__ __forceinline__ float zp_out() const { return 0.f; }
};

}  // namespace ar
