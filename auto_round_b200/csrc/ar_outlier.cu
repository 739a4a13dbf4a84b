// Outlier-suppressed block loss of enable_alg_ext -- SignRoundV2Quantizer._get_loss
// (auto_round/algorithms/quantization/sign_roundv2/quantizer.py:362-399):
//
//     diff  = |pred - ref|                      (bf16 tensor arithmetic)
//     top   = topk(diff.view(-1), max(1, numel // 1000)).indices       -> those elements are dropped
//     loss  = mean((|pred.float() - ref.float()| * token_mask * keep)^2)        over ALL elements
//
// The reference runs a global torch.topk over 67 M elements every iteration.  |diff| is a bf16 value, so it has at most
// 2^15 distinct bit patterns and its order is the order of those patterns: an exact selection needs only a 32768-bin
// histogram (pass 1), a suffix scan for the threshold pattern (pass 2, one block) and the loss/gradient pass (pass 3),
// which drops every element above the threshold and the first `need` ones found AT the threshold (torch.topk also breaks
// such ties by an implementation-defined choice).  All three are static launches: they live inside the captured CUDA graph.
#include "ar_common.cuh"

namespace ar {

namespace {

constexpr int kBins = 32768;

__device__ __forceinline__ void unpack8(const U4& r, float (&out)[8]) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = bf16_bits_to_f32((uint16_t)(u[i] & 0xffffu));
    out[2 * i + 1] = bf16_bits_to_f32((uint16_t)(u[i] >> 16));
  }
}

// bit pattern of |bf16(p - r)|: what `torch.abs(pred - ref)` holds for two bf16 tensors
__device__ __forceinline__ uint32_t absdiff_bits(float p, float r) { return (uint32_t)f32_to_bf16_bits(p - r) & 0x7fffu; }

// pass 1: block-private histogram in (opt-in) dynamic shared memory, flushed with one atomic per non-empty bin
__global__ void __launch_bounds__(1024) absdiff_hist_kernel(const U4* __restrict__ pred, const U4* __restrict__ ref,
                                                            int64_t n8, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float p[8], r[8];
    unpack8(pred[i], p);
    unpack8(ref[i], r);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sh[absdiff_bits(p[j], r[j])], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// pass 2: sel[0] = threshold pattern t with count(> t) < k <= count(>= t); sel[1] = k - count(> t) ("need": how many
// elements AT t are dropped); sel[2] = 0 (tie counter of pass 3).  Clears the histogram for the next iteration.
// Data parallel: `hist` holds `world` per-rank histograms [world, 32768] (all-gathered); the threshold comes from their sum
// (the top-k is global over the batch) and the `need` ties at the threshold are handed out in rank order, so that exactly
// k elements are dropped over all ranks: need_r = clamp(need - sum_{r' < r} ties_r', 0, ties_r).  `clear` (this rank's own
// histogram) is re-armed.
__global__ void __launch_bounds__(1024) topk_threshold_kernel(const uint32_t* __restrict__ hist, int world, int rank,
                                                              uint32_t* __restrict__ clear, unsigned long long k,
                                                              uint32_t* __restrict__ sel) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;                       // thread t owns bins [32 t, 32 t + 32)
  unsigned long long s = 0;
  uint32_t c[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    uint32_t v = 0;
    for (int r = 0; r < world; ++r) v += hist[(size_t)r * kBins + t * 32 + i];
    c[i] = v;
    s += v;
  }
  part[t] = s;
  __syncthreads();
  // inclusive suffix sum over the 1024 partials (Hillis-Steele, 10 steps)
  for (int off = 1; off < 1024; off <<= 1) {
    const unsigned long long add = (t + off < 1024) ? part[t + off] : 0ull;
    __syncthreads();
    part[t] += add;
    __syncthreads();
  }
  const unsigned long long above = (t + 1 < 1024) ? part[t + 1] : 0ull;   // elements in bins owned by higher threads
  if (above < k && part[t] >= k) {                 // the threshold bin is one of mine (exactly one thread qualifies)
    unsigned long long cum = above;
    for (int i = 31; i >= 0; --i) {
      if (cum + c[i] >= k) {
        unsigned long long need = k - cum;         // ties to drop over all ranks
        for (int r = 0; r < rank; ++r) {
          const unsigned long long mine = hist[(size_t)r * kBins + t * 32 + i];
          need = need > mine ? need - mine : 0ull;
        }
        const unsigned long long own = hist[(size_t)rank * kBins + t * 32 + i];
        sel[0] = (uint32_t)(t * 32 + i);
        sel[1] = (uint32_t)(need < own ? need : own);
        break;
      }
      cum += c[i];
    }
  }
  if (t == 0) {
    sel[2] = 0u;
    if (part[0] < k) { sel[0] = 0u; sel[1] = 0xffffffffu; }    // fewer than k elements in total: drop everything
  }
  __syncthreads();                                 // (world == 1: `clear` aliases `hist`; every read above is done)
#pragma unroll
  for (int i = 0; i < 32; ++i) clear[t * 32 + i] = 0u;
}

// pass 3: loss_sum += sum((|d| m keep)^2) (double, unnormalised); dpred = bf16 of autograd's chain
//   mean -> pow 2 -> * keep -> * token mask -> abs -> sub:   ((upstream / numel) * (2 x)) * sign(d),  x = |d| m keep
__global__ void __launch_bounds__(256) mse_outlier_kernel(const uint16_t* __restrict__ pred, const uint16_t* __restrict__ ref,
                                                          const uint8_t* __restrict__ mask, int64_t rows, int cols8,
                                                          float up_over_n, uint32_t* __restrict__ sel,
                                                          double* __restrict__ loss_sum, uint16_t* __restrict__ dpred) {
  const int64_t total = rows * (int64_t)cols8;
  const uint32_t thr = sel[0], need = sel[1];
  float local = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / cols8;
    const bool on = (mask == nullptr) || (mask[row] != 0);
    float p[8], r[8], g[8];
    unpack8(reinterpret_cast<const U4*>(pred)[i], p);
    unpack8(reinterpret_cast<const U4*>(ref)[i], r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // the selection ignores the token mask, exactly like the reference (topk runs on the unmasked difference)
      const uint32_t b = absdiff_bits(p[j], r[j]);
      bool keep = b < thr;
      if (b == thr) keep = !(atomicAdd(&sel[2], 1u) < need);
      const float d = p[j] - r[j];
      const float x = (on && keep) ? fabsf(d) : 0.f;
      local += x * x;
      const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
      g[j] = (up_over_n * (2.f * x)) * sg;
    }
    if (dpred) {
      uint32_t out[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = (uint32_t)f32_to_bf16_bits(g[2 * j]) | ((uint32_t)f32_to_bf16_bits(g[2 * j + 1]) << 16);
      reinterpret_cast<U4*>(dpred)[i] = U4{out[0], out[1], out[2], out[3]};
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ float warp_sums[8];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += (double)warp_sums[w];
    atomicAdd(loss_sum, s);
  }
}

}  // namespace

}  // namespace ar

using namespace ar;

extern "C" int ar_absdiff_hist(const void* pred, const void* ref, int64_t numel, uint32_t* hist, void* stream) {
  AR_REQUIRE(pred && ref && hist && numel > 0, AR_E_BADARG, "ar_absdiff_hist: bad arguments");
  AR_REQUIRE(numel % 8 == 0, AR_E_UNSUPPORTED, "ar_absdiff_hist: numel must be a multiple of 8");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(absdiff_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBins * 4);
    AR_REQUIRE(e == cudaSuccess, (int)e, "ar_absdiff_hist: cannot opt in to %d bytes of shared memory: %s", kBins * 4,
               cudaGetErrorString(e));
    configured = true;
  }
  const int64_t n8 = numel / 8;
  int64_t blocks = (n8 + 1023) / 1024;
  if (blocks > sm_count()) blocks = sm_count();        // 128 KB of shared memory per block: one block per SM
  if (blocks < 1) blocks = 1;
  absdiff_hist_kernel<<<(unsigned)blocks, 1024, kBins * 4, (cudaStream_t)stream>>>((const U4*)pred, (const U4*)ref, n8, hist);
  AR_CHECK_LAUNCH();
  return AR_OK;
}

extern "C" int ar_topk_threshold(uint32_t* hist, int64_t k, uint32_t* sel, void* stream) {
  AR_REQUIRE(hist && sel && k >= 1, AR_E_BADARG, "ar_topk_threshold: bad arguments");
  topk_threshold_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(hist, 1, 0, hist, (unsigned long long)k, sel);
  AR_CHECK_LAUNCH();
  return AR_OK;
}

extern "C" int ar_topk_threshold_ranks(const uint32_t* hist_all, int world, int rank, uint32_t* hist_local, int64_t k,
                                       uint32_t* sel, void* stream) {
  AR_REQUIRE(hist_all && hist_local && sel && k >= 1 && world >= 1 && rank >= 0 && rank < world, AR_E_BADARG,
             "ar_topk_threshold_ranks: bad arguments");
  topk_threshold_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(hist_all, world, rank, hist_local, (unsigned long long)k, sel);
  AR_CHECK_LAUNCH();
  return AR_OK;
}

extern "C" int ar_mse_outlier_fwd_bwd(const void* pred, const void* ref, const uint8_t* row_mask, int64_t rows, int64_t cols,
                                      int64_t numel_total, float upstream, uint32_t* sel, double* loss_sum, void* dpred,
                                      void* stream) {
  AR_REQUIRE(pred && ref && sel && loss_sum && rows > 0 && cols > 0, AR_E_BADARG, "ar_mse_outlier_fwd_bwd: bad arguments");
  AR_REQUIRE(cols % 8 == 0, AR_E_UNSUPPORTED, "ar_mse_outlier_fwd_bwd: cols must be a multiple of 8");
  const int64_t total = rows * (cols / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // mean backward: grad_output / numel; numel_total > 0: the mean runs over the GLOBAL batch of a data-parallel iteration
  const float up_over_n = upstream / (float)(numel_total > 0 ? numel_total : rows * cols);
  mse_outlier_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)pred, (const uint16_t*)ref, row_mask, rows,
                                                                         (int)(cols / 8), up_over_n, sel, loss_sum,
                                                                         (uint16_t*)dpred);
  AR_CHECK_LAUNCH();
  return AR_OK;
}
