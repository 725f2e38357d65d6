// letf_kernels.cuh — the fused path kernel (one thread = one trial at a time, persistent
// grid-stride over trials) and the registry that lets the C-ABI layer pick an instantiation.
#pragma once
#include <cuda_runtime.h>
#include "letf_core.cuh"
#include "letf_fast.cuh"
#include "letf_jfast.cuh"

namespace letf {

// trials per thread of the straight-line fast path (letf_fast.cuh); 1 = the build has none
template <typename R, int NA, int MARKET, bool TAPEIO>
LETF_HD constexpr int fast_upt() { return (sizeof(R) == 4 && NA < kAssets && MARKET == LETF_MARKET_NORMAL && !TAPEIO) ? 2 : 1; }

template <typename R, int NS, int NA, int MARKET, bool TAPEIO, uint32_t TM, int LIST, int AM, int BLK, int MINB>
__global__ void __launch_bounds__(BLK, MINB) letf_path_kernel(const __grid_constant__ KArgs<R, NS, NA> ka) {
  const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int UPT = fast_upt<R, NA, MARKET, TAPEIO>();
  if constexpr (UPT > 1) {
    if (fast_pattern_compiled<NA>(ka.p.fast_layers) && ka.rng_mode == LETF_RNG_PHILOX) {   // grid sized for ceil(n / UPT) threads
      for (int64_t j = lane; j < ka.n_trials; j += (int64_t)UPT * ka.nthreads) {
        int64_t trial[UPT], col[UPT];
        bool live[UPT];
#pragma unroll
        for (int u = 0; u < UPT; u++) {
          col[u] = j + (int64_t)u * ka.nthreads;
          live[u] = col[u] < ka.n_trials;
          if (!live[u]) col[u] = j;                   // padding lane recomputes trial j and stores nothing
          trial[u] = ka.first_trial + col[u];
        }
        fast_layers_run<NS, NA, UPT>(ka, trial, col, live);
      }
      return;
    }
  }
  for (int64_t j = lane; j < ka.n_trials; j += ka.nthreads) {
    Trial<R, NS, NA, MARKET, TAPEIO, TM, LIST, AM> t(ka, (TAPEIO && ka.trial_ids) ? ka.trial_ids[j] : ka.first_trial + j, j, lane);
    t.run();
  }
}

// one antithetic pair per thread (letf_jfast.cuh): work item j = pair (first_trial >> 1) + j; a twin outside
// [first_trial, first_trial + n_trials) is computed and not stored
template <int NS, int NA, int NACT, int K10, bool EMIT, int BLK, int MINB>
__global__ void __launch_bounds__(BLK, MINB) letf_jpair_kernel(const __grid_constant__ KArgs<float, NS, NA> ka) {
  // both regimes' constants in shared memory: every lane reads them with its own regime index (letf_jfast.cuh)
  __shared__ DevFJ<float> s_fj[2];
  __shared__ DevJoint<float> s_jt[2];
  for (int i = threadIdx.x; i < (int)(sizeof(s_fj) / 4); i += BLK) ((uint32_t*)s_fj)[i] = ((const uint32_t*)ka.p.fj)[i];
  for (int i = threadIdx.x; i < (int)(sizeof(s_jt) / 4); i += BLK) ((uint32_t*)s_jt)[i] = ((const uint32_t*)ka.p.joint)[i];
  __syncthreads();
  const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p0 = ka.first_trial >> 1, p1 = (ka.first_trial + ka.n_trials - 1) >> 1;
  for (int64_t pj = p0 + lane; pj <= p1; pj += ka.nthreads) JointPair<NS, NA, NACT, K10, EMIT>::run(ka, pj, lane, s_fj, s_jt);
}

// type-erased launcher table -----------------------------------------------------------------
struct KernelEntry {
  int precision;      // LETF_PREC_*
  int ns;             // strategy slots compiled in
  int na;             // ETF series slots compiled in (6 = all, fewer = compacted lean build)
  int market;         // LETF_MARKET_* (3 = caller-supplied series)
  int tapeio;         // 1: can read / record tapes and keeps the trade journal
  uint32_t tm;        // strategy types compiled in
  int list;           // typed slot list (slot_list_type), 0 = dynamic dispatch
  int am;             // compile-time asset mask, 0 = run time
  int block;          // threads per block
  int upt;            // trials per thread when DevParams::fast_layers applies (letf_fast.cuh), else 1
  int fj;             // DevParams::fast_joint code this entry serves (antithetic-pair kernel, letf_jfast.cuh), 0 = any shape
  const char* name;
  cudaError_t (*launch)(const void* kargs, int grid, cudaStream_t st);
  cudaError_t (*attrs)(cudaFuncAttributes* a, int* max_blocks_per_sm);
  size_t kargs_size;
};

void register_kernel(const KernelEntry& e);
int kernel_count();
const KernelEntry* kernel_at(int i);

template <typename R, int NS, int NA, int MARKET, bool TAPEIO, uint32_t TM, int LIST, int AM, int BLK, int MINB>
struct KernelImpl {
  // record mode stages kStage days x kRecValues values per thread through dynamic shared memory (letf_core.cuh)
  static constexpr size_t kSmem = MARKET == kMarketFSeries
      ? (size_t)kStage * kRecValues<__builtin_popcount((unsigned)AM & 0x3Fu) + 2>() * BLK * sizeof(float) : 0;
  static cudaError_t launch(const void* kargs, int grid, cudaStream_t st) {
    if (kSmem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(letf_path_kernel<R, NS, NA, MARKET, TAPEIO, TM, LIST, AM, BLK, MINB>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
      if (e != cudaSuccess) return e;
    }
    letf_path_kernel<R, NS, NA, MARKET, TAPEIO, TM, LIST, AM, BLK, MINB><<<grid, BLK, kSmem, st>>>(*(const KArgs<R, NS, NA>*)kargs);
    return cudaGetLastError();
  }
  static cudaError_t attrs(cudaFuncAttributes* a, int* mb) {
    cudaError_t e = cudaFuncGetAttributes(a, letf_path_kernel<R, NS, NA, MARKET, TAPEIO, TM, LIST, AM, BLK, MINB>);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(mb, letf_path_kernel<R, NS, NA, MARKET, TAPEIO, TM, LIST, AM, BLK, MINB>, BLK, kSmem);
  }
};

#define LETF_INSTANTIATE(ID, R, PREC, NS, NA, MARKET, TAPEIO, TM, LIST, AM, BLK, MINB)                   \
  namespace {                                                                                            \
  struct Reg_##ID {                                                                                      \
    Reg_##ID() {                                                                                         \
      typedef letf::KernelImpl<R, NS, NA, MARKET, TAPEIO, TM, LIST, AM, BLK, MINB> K;                    \
      letf::KernelEntry e = {PREC, NS, NA, MARKET, TAPEIO ? 1 : 0, TM, LIST, AM, BLK,                    \
                             letf::fast_upt<R, NA, MARKET, TAPEIO>(), 0, #ID, &K::launch,                \
                             &K::attrs,                                                                  \
                             sizeof(letf::KArgs<R, NS, NA>)};                                                \
      letf::register_kernel(e);                                                                          \
    }                                                                                                    \
  } reg_##ID;                                                                                            \
  }

template <int NS, int NA, int NACT, int K10, bool EMIT, int BLK, int MINB>
struct JPairImpl {
  static cudaError_t launch(const void* kargs, int grid, cudaStream_t st) {
    letf_jpair_kernel<NS, NA, NACT, K10, EMIT, BLK, MINB><<<grid, BLK, 0, st>>>(*(const KArgs<float, NS, NA>*)kargs);
    return cudaGetLastError();
  }
  static cudaError_t attrs(cudaFuncAttributes* a, int* mb) {
    cudaError_t e = cudaFuncGetAttributes(a, letf_jpair_kernel<NS, NA, NACT, K10, EMIT, BLK, MINB>);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(mb, letf_jpair_kernel<NS, NA, NACT, K10, EMIT, BLK, MINB>, BLK, 0);
  }
};
constexpr int kSplitProducer = 0x10000;      // KernelEntry::fj flag: EMIT build (producer of the split strategy pipeline)

#define LETF_INSTANTIATE_JPAIR(ID, NS, NA, NACT, K10, EMIT, BLK, MINB)                                   \
  namespace {                                                                                            \
  struct Reg_##ID {                                                                                      \
    Reg_##ID() {                                                                                         \
      typedef letf::JPairImpl<NS, NA, NACT, K10, EMIT, BLK, MINB> K;                                     \
      letf::KernelEntry e = {LETF_PREC_F32, NS, NA, LETF_MARKET_JOINT_T, 0, letf::TM_BENCH, 0, 0, BLK, 2, \
                             letf::fast_joint_code(NACT, K10) | (EMIT ? letf::kSplitProducer : 0), #ID,  \
                             &K::launch, &K::attrs,                                                      \
                             sizeof(letf::KArgs<float, NS, NA>)};                                        \
      letf::register_kernel(e);                                                                          \
    }                                                                                                    \
  } reg_##ID;                                                                                            \
  }

constexpr uint32_t TM_ALL = (1u << LETF_ST__COUNT) - 1u;
constexpr uint32_t TM_BENCH = (1u << LETF_ST_BENCHMARK) | (1u << LETF_ST_BENCHMARK_TLT);
// BASELINE.json config 4: benchmarks + SMA switch + 60/40 portfolio + threshold-rebalanced S19
constexpr uint32_t TM_C4 = TM_BENCH | (1u << LETF_ST_SMA) | (1u << LETF_ST_SMA_BAND) | (1u << LETF_ST_PORTFOLIO) |
                           (1u << LETF_ST_CONVICTION);

}  // namespace letf
