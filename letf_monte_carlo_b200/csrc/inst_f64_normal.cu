// one instantiation of the fused path kernel per translation unit (parallel builds)
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f64_normal, double, LETF_PREC_F64, 32, 6, 0, true, letf::TM_ALL, 0, 0, 128, 1)
