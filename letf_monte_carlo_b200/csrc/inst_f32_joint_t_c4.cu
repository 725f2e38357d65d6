// one instantiation of the fused path kernel per translation unit (parallel builds)
#include "letf_kernels.cuh"
// BASELINE config 4: S1 S2 S4 S5 S6 S19, one compile-time rule per slot, series TQQQ / TMF / SPY
LETF_INSTANTIATE(f32_joint_t_c4, float, LETF_PREC_F32, 8, 6, 2, false, letf::slot_list_mask(1), 1, 25, 64, 6)
