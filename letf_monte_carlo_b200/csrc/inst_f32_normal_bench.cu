// one instantiation of the fused path kernel per translation unit (parallel builds)
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_normal_bench, float, LETF_PREC_F32, 4, 3, 0, false, letf::TM_BENCH, 0, 0, 64, 6)
