// one instantiation of the fused path kernel per translation unit (parallel builds)
// the default 19-strategy list as typed slots, tracking-error degrees of freedom read at run time (asset mask 29 | bit 6)
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_joint_t_sweep19_tedf, float, LETF_PREC_F32, 32, 6, 2, false, letf::slot_list_mask(2), 2, 29 | 0x40, 128, 1)
