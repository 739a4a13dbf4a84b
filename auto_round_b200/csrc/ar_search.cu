// Optimized-RTN / alg_ext scale searches and the importance-matrix accumulator.
//
//   search_int   auto_round/data_type/int.py:24-86 (search_scales) + the q_scale_thresh clip and the bf16 qdq of
//                opt_rtn_int_sym (int.py:89-122).  The reference runs 200 (180 for 2 bit) full passes over W, one per
//                candidate; here a group stays in registers for the whole grid, W is read once.
//   search_nv    auto_round/data_type/nvfp.py:331-385 (search_nvfp4_scale): coefficient in {1, 0.50 .. 1.51}
//   search_mx    auto_round/data_type/mxfp.py:103-169 (search_mx_scale):    coefficient in {1, 0.5, 2}
//   imatrix      algorithms/quantization/rtn/quantizer.py:86-105 (collect_imatrix): sum over tokens of x^2 per channel
//
// Layout: a group of G consecutive K-elements is owned by LPG = min(32, G) adjacent lanes, G/LPG elements per lane;
// the per-candidate loss is a lane-local sum followed by an xor-shuffle tree.  ALU-bound (candidates x weights), not
// HBM-bound: W4 g128 Llama-3-8B block = 218 M weights x 201 candidates.
// Selection uses `loss < best` exactly as the reference; the fp32 loss sum is taken in a different order than torch's
// (as the reference's own CPU and CUDA runs differ), so near-tied candidates can resolve differently.
#include "ar_qdq_math.cuh"

namespace ar {

namespace {

constexpr int kSearchThreads = 256;

template <int LPG>
__device__ __forceinline__ float lanes_sum(float x) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
template <int LPG>
__device__ __forceinline__ float lanes_max(float x) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

struct SearchArgs {
  const uint16_t* w;     // bf16 [N, K]
  const float* qw;       // importance: [K] (row_stride 0) or [N, kpad] (row_stride kpad); nullptr -> 1
  long long qw_stride;
  const float* coef;     // candidate table, base candidate first
  int ncand;
  const float* gscale;   // NVFP4 per-tensor global scale
  int n, k, kpad, bits;
  float thr;
  float* out;            // [N * kpad / G]
  uint16_t* wq;          // optional bf16 [N, K] (int search only)
};

template <int G>
struct Geo {
  static constexpr int LPG = G < 32 ? G : 32;
  static constexpr int EPL = G / LPG;
  static constexpr int GPW = 32 / LPG;   // groups per warp
};

// loads this lane's EPL elements of group `gidx` (zero beyond K, as the reference zero-pads) and their loss weights
template <int G>
__device__ __forceinline__ void load_group(const SearchArgs& a, long long gidx, int sub, float (&x)[Geo<G>::EPL],
                                           float (&q)[Geo<G>::EPL], int& row, int& k0) {
  constexpr int EPL = Geo<G>::EPL;
  const int gpr = a.kpad / G;
  row = (int)(gidx / gpr);
  k0 = (int)(gidx % gpr) * G + sub * EPL;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int kk = k0 + i;
    x[i] = (kk < a.k) ? bf16_bits_to_f32(a.w[(long long)row * a.k + kk]) : 0.f;
    if (a.qw == nullptr) q[i] = 1.f;
    else if (a.qw_stride == 0) q[i] = (kk < a.k) ? a.qw[kk] : 1e-5f;          // pad value of the reference (int.py:111)
    else q[i] = a.qw[(long long)row * a.qw_stride + kk];
  }
}

// torch get_reciprocal (utils/common.py:903-922) on a bf16 tensor: |x| >= bf16(1e-30) ? bf16(1/x) : 0
__device__ __forceinline__ float recip_bf16(float x) {
  const float eps = bf16_round(1e-30f);
  return (fabsf(x) >= eps) ? bf16_round(1.f / x) : 0.f;
}

template <int G>
__global__ void __launch_bounds__(kSearchThreads) search_int_kernel(SearchArgs a, long long total_groups) {
  constexpr int LPG = Geo<G>::LPG, EPL = Geo<G>::EPL, GPW = Geo<G>::GPW;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * kSearchThreads + threadIdx.x) >> 5;
  const long long gidx = warp * GPW + lane / LPG;
  const int sub = lane % LPG;
  const bool valid = gidx < total_groups;
  const long long gsafe = valid ? gidx : total_groups - 1;
  float x[EPL], q[EPL];
  int row, k0;
  load_group<G>(a, gsafe, sub, x, q, row, k0);

  // signed value at the FIRST argmax of |x| (torch.argmax)
  float am = -1.f, gv = 0.f;
  int ai = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const float v = fabsf(x[i]);
    if (v > am) { am = v; ai = sub * EPL + i; gv = x[i]; }
  }
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, am, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ai, o);
    const float ov = __shfl_xor_sync(0xffffffffu, gv, o);
    if (om > am || (om == am && oi < ai)) { am = om; ai = oi; gv = ov; }
  }
  const float r = recip_bf16(gv);
  const float nmax = (float)(1 << (a.bits - 1));
  float best = 0.f, best_sc = 0.f;
  for (int c = 0; c < a.ncand; ++c) {
    const float isc = bf16_round(a.coef[c] * r);
    const float sc = recip_bf16(isc);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const float L = clampf(rintf(bf16_round(isc * x[i])), -nmax, nmax - 1.f);
      const float d = bf16_round(bf16_round(sc * L) - x[i]);
      part += (d * d) * q[i];
    }
    const float loss = lanes_sum<LPG>(part);
    if (c == 0 || loss < best) { best = loss; best_sc = sc; }
  }
  const float thr = bf16_round(a.thr);                                   // clamp runs in the bf16 tensor's dtype
  const float s = (best_sc < 0.f) ? fminf(best_sc, -thr) : fmaxf(best_sc, thr);
  if (!valid) return;
  if (sub == 0) a.out[gidx] = s;
  if (a.wq != nullptr) {
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int kk = k0 + i;
      if (kk < a.k) {
        const float L = clampf(rintf(bf16_round(x[i] / s)), -nmax, nmax - 1.f);
        a.wq[(long long)row * a.k + kk] = f32_to_bf16_bits(L * s);
      }
    }
  }
}

// NVFP4 (G = 16) and MXFP4 (G = 32) coefficient searches share one body: candidate c scales the group amax
template <class Ctx, int G>
__global__ void __launch_bounds__(kSearchThreads) search_fp4_kernel(SearchArgs a, long long total_groups) {
  constexpr int LPG = Geo<G>::LPG, EPL = Geo<G>::EPL, GPW = Geo<G>::GPW;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * kSearchThreads + threadIdx.x) >> 5;
  const long long gidx = warp * GPW + lane / LPG;
  const int sub = lane % LPG;
  const bool valid = gidx < total_groups;
  const long long gsafe = valid ? gidx : total_groups - 1;
  float x[EPL], q[EPL];
  int row, k0;
  load_group<G>(a, gsafe, sub, x, q, row, k0);
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) m = fmaxf(m, fabsf(x[i]));
  GroupIn gi;
  gi.wmax = lanes_max<LPG>(m);
  gi.wmin = 0.f;
  gi.mn = 1.f;
  gi.thr = a.thr;
  gi.gscale = a.gscale ? *a.gscale : 0.f;
  Ctx ctx;
  ctx.init(a.bits);
  float best = 0.f, best_c = 1.f;
  for (int c = 0; c < a.ncand; ++c) {
    gi.mx = a.coef[c];
    ctx.setup(gi);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const float d = ctx.fwd(x[i], 0.f) - x[i];
      part += (d * d) * q[i];
    }
    const float loss = lanes_sum<LPG>(part);
    if (c == 0 || loss < best) { best = loss; best_c = a.coef[c]; }
  }
  if (valid && sub == 0) a.out[gidx] = best_c;
}

// imatrix[k] += sum_rows x[row, k]^2.  4 columns per thread (8-byte loads), grid.y row slabs, fp32 atomics.
__global__ void __launch_bounds__(128) imatrix_kernel(const uint16_t* x, long long rows, int k, long long rows_per_slab,
                                                      float* imatrix) {
  const int c0 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c0 >= k) return;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  const long long r1 = (r0 + rows_per_slab < rows) ? r0 + rows_per_slab : rows;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = (k % 4 == 0);
  for (long long r = r0; r < r1; ++r) {
    const uint16_t* p = x + r * k + c0;
    if (vec) {
      const U2 u = *reinterpret_cast<const U2*>(p);
      const float v0 = bf16_bits_to_f32((uint16_t)(u.x & 0xffffu)), v1 = bf16_bits_to_f32((uint16_t)(u.x >> 16));
      const float v2 = bf16_bits_to_f32((uint16_t)(u.y & 0xffffu)), v3 = bf16_bits_to_f32((uint16_t)(u.y >> 16));
      acc[0] += v0 * v0; acc[1] += v1 * v1; acc[2] += v2 * v2; acc[3] += v3 * v3;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + i < k) { const float v = bf16_bits_to_f32(p[i]); acc[i] += v * v; }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (c0 + i < k) atomicAdd(imatrix + c0 + i, acc[i]);
}

int check_common(const void* w, const float* coef, int ncand, const ar_qspec* s, const float* out) {
  AR_REQUIRE(w && s && out, AR_E_BADARG, "search: null pointer");
  AR_REQUIRE(s->n > 0 && s->k > 0 && s->group_size > 0, AR_E_BADARG, "search: bad shape n=%d k=%d g=%d", s->n, s->k,
             s->group_size);
  AR_REQUIRE(ncand >= 1 && ncand <= 4096 && coef, AR_E_BADARG, "search: candidate table (ncand=%d)", ncand);
  return 0;
}

}  // namespace

}  // namespace ar

using namespace ar;

extern "C" int ar_search_scale_int(const void* w, const float* qw, long long qw_row_stride, const float* coef, int ncand,
                                   const ar_qspec* s, float* scale, void* wq, void* stream) {
  if (int e = check_common(w, coef, ncand, s, scale)) return e;
  AR_REQUIRE(s->dtype == AR_DT_INT_SYM, AR_E_BADARG, "ar_search_scale_int: int_sym only (int.py:89-122)");
  AR_REQUIRE(s->bits >= 2 && s->bits <= 8, AR_E_BADARG, "ar_search_scale_int: bits=%d", s->bits);
  const int g = s->group_size;
  SearchArgs a{};
  a.w = (const uint16_t*)w; a.qw = qw; a.qw_stride = qw_row_stride; a.coef = coef; a.ncand = ncand; a.gscale = nullptr;
  a.n = s->n; a.k = s->k; a.kpad = (s->k + g - 1) / g * g; a.bits = s->bits; a.thr = s->q_scale_thresh;
  a.out = scale; a.wq = (uint16_t*)wq;
  AR_REQUIRE(qw == nullptr || qw_row_stride == 0 || qw_row_stride >= a.kpad, AR_E_BADARG, "search: qw_row_stride");
  const long long total = (long long)a.n * (a.kpad / g);
  cudaStream_t st = (cudaStream_t)stream;
#define AR_LAUNCH_INT(GG)                                                                              \
  case GG: {                                                                                           \
    const long long warps = (total + Geo<GG>::GPW - 1) / Geo<GG>::GPW;                                 \
    const long long blocks = (warps * 32 + kSearchThreads - 1) / kSearchThreads;                       \
    search_int_kernel<GG><<<(unsigned)blocks, kSearchThreads, 0, st>>>(a, total);                      \
  } break;
  switch (g) {
    AR_LAUNCH_INT(16) AR_LAUNCH_INT(32) AR_LAUNCH_INT(64) AR_LAUNCH_INT(128) AR_LAUNCH_INT(256)
    default: AR_REQUIRE(false, AR_E_UNSUPPORTED, "ar_search_scale_int: group_size %d (16/32/64/128/256)", g);
  }
#undef AR_LAUNCH_INT
  AR_CHECK_LAUNCH();
  return 0;
}

template <class Ctx, int G>
static int search_fp4(const void* w, const float* qw, long long qw_row_stride, const float* gscale, const float* coef,
                      int ncand, const ar_qspec* s, float* out, void* stream) {
  if (int e = check_common(w, coef, ncand, s, out)) return e;
  AR_REQUIRE(s->group_size == G, AR_E_UNSUPPORTED, "fp4 scale search: group_size %d (expected %d)", s->group_size, G);
  SearchArgs a{};
  a.w = (const uint16_t*)w; a.qw = qw; a.qw_stride = qw_row_stride; a.coef = coef; a.ncand = ncand; a.gscale = gscale;
  a.n = s->n; a.k = s->k; a.kpad = (s->k + G - 1) / G * G; a.bits = 4; a.thr = s->q_scale_thresh;
  a.out = out; a.wq = nullptr;
  AR_REQUIRE(qw == nullptr || qw_row_stride == 0 || qw_row_stride >= a.kpad, AR_E_BADARG, "search: qw_row_stride");
  const long long total = (long long)a.n * (a.kpad / G);
  const long long warps = (total + Geo<G>::GPW - 1) / Geo<G>::GPW;
  const long long blocks = (warps * 32 + kSearchThreads - 1) / kSearchThreads;
  search_fp4_kernel<Ctx, G><<<(unsigned)blocks, kSearchThreads, 0, (cudaStream_t)stream>>>(a, total);
  AR_CHECK_LAUNCH();
  return 0;
}

extern "C" int ar_search_scale_nv(const void* w, const float* qw, long long qw_row_stride, const float* gscale,
                                  const float* coef, int ncand, const ar_qspec* s, float* coeff_out, void* stream) {
  AR_REQUIRE(gscale, AR_E_BADARG, "ar_search_scale_nv: gscale (448*6/amax of THIS tensor, nvfp.py:335) is required");
  return search_fp4<NvFp4, 16>(w, qw, qw_row_stride, gscale, coef, ncand, s, coeff_out, stream);
}

extern "C" int ar_search_scale_mx(const void* w, const float* qw, long long qw_row_stride, const float* coef, int ncand,
                                  const ar_qspec* s, float* coeff_out, void* stream) {
  return search_fp4<MxFp4, 32>(w, qw, qw_row_stride, nullptr, coef, ncand, s, coeff_out, stream);
}

extern "C" int ar_imatrix_accum(const void* x, long long rows, int k, float* imatrix, void* stream) {
  AR_REQUIRE(x && imatrix && rows >= 0 && k > 0, AR_E_BADARG, "ar_imatrix_accum: bad arguments");
  if (rows == 0) return 0;
  const int gx = (k + 511) / 512;
  int slabs = (sm_count() * 8 + gx - 1) / gx;
  if (slabs > rows) slabs = (int)rows;
  if (slabs < 1) slabs = 1;
  const long long per = (rows + slabs - 1) / slabs;
  imatrix_kernel<<<dim3(gx, (unsigned)((rows + per - 1) / per)), 128, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)x, rows, k, per, imatrix);
  AR_CHECK_LAUNCH();
  return 0;
}
