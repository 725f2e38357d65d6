// one instantiation of the fused path kernel per translation unit (parallel builds)
// BASELINE config 4 list as typed slots, tracking-error degrees of freedom read at run time (asset mask 25 | bit 6):
// what a calibrated tracking_residual_model gets instead of the generic 32-slot build
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_joint_t_c4_tedf, float, LETF_PREC_F32, 8, 6, 2, false, letf::slot_list_mask(1), 1, 25 | 0x40, 64, 6)
