// one instantiation of the fused path kernel per translation unit (parallel builds)
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_joint_t_io, float, LETF_PREC_F32, 32, 6, 2, true, letf::TM_ALL, 0, 0, 128, 1)
