// kernel + launcher of the materialised-path record pass (letf_series_prep.cuh)
#include <cuda_runtime.h>
#include "letf_series_prep.cuh"

namespace letf {

__global__ void __launch_bounds__(128) letf_series_records_kernel(const SeriesPrepArgs a) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < a.n) series_prep_path(a, j);
}

cudaError_t launch_series_records(const SeriesPrepArgs& a, cudaStream_t st) {
  if (a.n <= 0) return cudaSuccess;
  const int block = 128;
  const int64_t grid = (a.n + block - 1) / block;
  letf_series_records_kernel<<<(unsigned)grid, block, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace letf
