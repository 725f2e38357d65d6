// one instantiation of the fused path kernel per translation unit (parallel builds)
// the reference's default 19-strategy sweep with one compile-time rule per slot
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_joint_t_sweep19, float, LETF_PREC_F32, 32, 6, 2, false, letf::slot_list_mask(2), 2, 29, 128, 1)
