// letf_reduce.cu — summary statistics of the per-trial outputs: the quantities the reference's
// reporting layer derives from Final_Wealth / Max_DD / Trades_Per_Year (letf/reporting.py:358-376,
// 320-325, 541-555), in a form that is sum/min/max-reducible across GPUs:
//   per (strategy, metric) a 4096-bin histogram  (metric 0: log10(W/W0), 1: CAGR, 2: max drawdown)
//   sum x, sum x^2, min, max;  counters: valid trials, wins vs the benchmark strategy BY TRIAL INDEX,
//   total trades, flagged trials.
// Each block owns one strategy and a slice of trials; bins are accumulated in a shared-memory
// histogram, moments with warp shuffles; one flush of non-zero bins per block to HBM.
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/letf_b200.h"

namespace letf {

constexpr int kRBlock = 256;

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

struct ReduceArgs {
  LetfOutputs out;
  LetfSummary sm;
  int64_t n;
  double w0, inv_years;
};

__global__ void __launch_bounds__(kRBlock) letf_reduce_kernel(const __grid_constant__ ReduceArgs ra) {
  extern __shared__ unsigned int hist_dyn[];
  unsigned int (*hist)[LETF_HIST_BINS] = (unsigned int (*)[LETF_HIST_BINS])hist_dyn;
  __shared__ double red[kRBlock / 32][LETF_N_METRICS][4];
  __shared__ unsigned long long cred[kRBlock / 32][4];
  const int s = blockIdx.y;
  const int64_t N = ra.out.n_trials, n = ra.n;
  for (int i = threadIdx.x; i < LETF_N_METRICS * LETF_HIST_BINS; i += kRBlock) (&hist[0][0])[i] = 0u;
  __syncthreads();
  double sx[3] = {0, 0, 0}, sxx[3] = {0, 0, 0}, mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  unsigned long long c_valid = 0, c_win = 0, c_trades = 0, c_flag = 0;
  const double* W = ra.out.final_wealth + (int64_t)s * N;
  const double* D = ra.out.max_dd + (int64_t)s * N;
  const int32_t* TR = ra.out.num_trades + (int64_t)s * N;
  const double* WB = ra.sm.bench_strategy >= 0 ? ra.out.final_wealth + (int64_t)ra.sm.bench_strategy * N : nullptr;
  for (int64_t j = (int64_t)blockIdx.x * kRBlock + threadIdx.x; j < n; j += (int64_t)gridDim.x * kRBlock) {
    double w = W[j];
    unsigned int fl = ra.out.flags ? ra.out.flags[j] : 0u;
    if (fl & (LETF_FLAG_NONFINITE_RET | LETF_FLAG_BAD_VIX)) { c_flag++; continue; }   // reference raises on these
    if (fl) c_flag++;
    double ratio = w / ra.w0;
    double m[3];
    m[0] = log10(fmax(ratio, 1e-300));
    m[1] = pow(fmax(ratio, 0.0), ra.inv_years) - 1.0;      // reporting.py:320-325
    m[2] = D[j];
    c_valid++;
    c_trades += (unsigned long long)TR[j];
    if (WB && w > WB[j]) c_win++;                          // reporting.py:541-555, by trial index
#pragma unroll
    for (int q = 0; q < 3; q++) {
      double v = m[q];
      sx[q] += v; sxx[q] += v * v; mn[q] = fmin(mn[q], v); mx[q] = fmax(mx[q], v);
      double t = (v - ra.sm.hist_lo[q]) / (ra.sm.hist_hi[q] - ra.sm.hist_lo[q]) * (double)LETF_HIST_BINS;
      int b = t < 0 ? 0 : (t >= (double)LETF_HIST_BINS ? LETF_HIST_BINS - 1 : (int)t);
      atomicAdd(&hist[q][b], 1u);
    }
  }
  const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 3; q++) {
    double a = warp_sum(sx[q]), b = warp_sum(sxx[q]), c = warp_min(mn[q]), d = warp_max(mx[q]);
    if (ln == 0) { red[warp][q][0] = a; red[warp][q][1] = b; red[warp][q][2] = c; red[warp][q][3] = d; }
  }
  {
    unsigned long long a = warp_sum_u(c_valid), b = warp_sum_u(c_win), c = warp_sum_u(c_trades), d = warp_sum_u(c_flag);
    if (ln == 0) { cred[warp][0] = a; cred[warp][1] = b; cred[warp][2] = c; cred[warp][3] = d; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int q = threadIdx.x;
    double a = 0, b = 0, c = INFINITY, d = -INFINITY;
    for (int w = 0; w < kRBlock / 32; w++) { a += red[w][q][0]; b += red[w][q][1]; c = fmin(c, red[w][q][2]); d = fmax(d, red[w][q][3]); }
    atomicAdd(&ra.sm.moments[((int64_t)s * 3 + q) * 2 + 0], a);
    atomicAdd(&ra.sm.moments[((int64_t)s * 3 + q) * 2 + 1], b);
    atomic_min_double(&ra.sm.extrema[((int64_t)s * 3 + q) * 2 + 0], c);
    atomic_max_double(&ra.sm.extrema[((int64_t)s * 3 + q) * 2 + 1], d);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 36) {
    int q = threadIdx.x - 32;
    unsigned long long a = 0;
    for (int w = 0; w < kRBlock / 32; w++) a += cred[w][q];
    atomicAdd(&ra.sm.counters[(int64_t)s * 4 + q], a);
  }
  for (int i = threadIdx.x; i < LETF_N_METRICS * LETF_HIST_BINS; i += kRBlock) {
    unsigned int c = (&hist[0][0])[i];
    if (c) atomicAdd(&ra.sm.hist[(int64_t)s * LETF_N_METRICS * LETF_HIST_BINS + i], (unsigned long long)c);
  }
}

cudaError_t launch_reduce(const LetfOutputs& out, const LetfSummary& sm, int64_t n, int n_strat, double w0,
                          double years, int sm_count, cudaStream_t st) {
  if (n <= 0 || n_strat <= 0) return cudaSuccess;
  ReduceArgs ra;
  ra.out = out; ra.sm = sm; ra.n = n; ra.w0 = w0; ra.inv_years = years > 0 ? 1.0 / years : 0.0;
  int64_t want = (n + kRBlock - 1) / kRBlock;
  int per = sm_count * 2 / (n_strat > 0 ? n_strat : 1);
  if (per < 1) per = 1;
  int gx = (int)(want < per ? want : per);
  dim3 grid(gx, n_strat);
  const int smem = LETF_N_METRICS * LETF_HIST_BINS * (int)sizeof(unsigned int);
  // the opt-in is a per-DEVICE attribute of the function and cheap to set: set it on every launch (a process-wide flag
  // left a second device driven from the same process without it)
  {
    cudaError_t e = cudaFuncSetAttribute(letf_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  letf_reduce_kernel<<<grid, kRBlock, smem, st>>>(ra);
  return cudaGetLastError();
}

}  // namespace letf

namespace letf {
__global__ void letf_summary_reset_kernel(LetfSummary sm, int n_strat) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t nh = (int64_t)n_strat * LETF_N_METRICS * LETF_HIST_BINS;
  if (i < nh) sm.hist[i] = 0ull;
  if (i < (int64_t)n_strat * LETF_N_METRICS * 2) {
    sm.moments[i] = 0.0;
    sm.extrema[i] = (i & 1) ? -INFINITY : INFINITY;
  }
  if (i < (int64_t)n_strat * 4) sm.counters[i] = 0ull;
}
cudaError_t launch_summary_reset(const LetfSummary& sm, int n_strat, cudaStream_t st) {
  int64_t nh = (int64_t)n_strat * LETF_N_METRICS * LETF_HIST_BINS;
  letf_summary_reset_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(sm, n_strat);
  return cudaGetLastError();
}
}  // namespace letf

// ---- exchange step: fold `world` packed summaries (one per rank, gathered over NCCL) into one --------------------------
namespace letf {
__global__ void letf_summary_combine_kernel(const unsigned long long* __restrict__ g, int world, int64_t words, int64_t n_int,
                                            int64_t n_mom, unsigned long long* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= words) return;
  if (i < n_int) {                         // histogram bins and counters: integer sums
    unsigned long long a = 0;
    for (int r = 0; r < world; r++) a += g[(int64_t)r * words + i];
    dst[i] = a;
  } else if (i < n_int + n_mom) {          // sum x, sum x^2 (rank order is fixed: the result is the same on every rank)
    double a = 0.0;
    for (int r = 0; r < world; r++) a += __longlong_as_double((long long)g[(int64_t)r * words + i]);
    dst[i] = (unsigned long long)__double_as_longlong(a);
  } else {                                 // extrema [.., 2]: even word = min, odd word = max
    const bool is_max = ((i - n_int - n_mom) & 1) != 0;
    double a = is_max ? -INFINITY : INFINITY;
    for (int r = 0; r < world; r++) {
      const double v = __longlong_as_double((long long)g[(int64_t)r * words + i]);
      a = is_max ? fmax(a, v) : fmin(a, v);
    }
    dst[i] = (unsigned long long)__double_as_longlong(a);
  }
}
cudaError_t launch_summary_combine(const void* gathered, int world, int n_strat, void* dst, cudaStream_t st) {
  const int64_t n_int = (int64_t)n_strat * (LETF_N_METRICS * LETF_HIST_BINS + 4), n_mom = (int64_t)n_strat * LETF_N_METRICS * 2;
  const int64_t words = n_int + 2 * n_mom;
  letf_summary_combine_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>((const unsigned long long*)gathered, world, words,
                                                                                n_int, n_mom, (unsigned long long*)dst);
  return cudaGetLastError();
}
}  // namespace letf
