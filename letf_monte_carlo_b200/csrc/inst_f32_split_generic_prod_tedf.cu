// one instantiation of the fused path kernel per translation unit (parallel builds)
// producer of the generic split strategy pipeline for a calibrated tracking_residual_model (per-series dof, tanh cap):
// the tracking-error sampler dispatches at run time
#include "letf_kernels.cuh"
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_all_tedf, 32, 6, 6, 0x80, true, 128, 2)
