// one instantiation of the fused path kernel per translation unit (parallel builds)
// rule kernel of the split strategy pipeline for any SUBSET of the config-4 list (S1 S2 S4 S5 S6 S19 on TQQQ / TMF / SPY;
// slot list 4 = list 1 with routed slots): the cheapest kernel for the common "SMA switch against buy&hold" requests
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_rules_sub6, float, LETF_PREC_F32, 8, 6, 5, false, letf::slot_list_mask(4), 4, 25, 128, 3)
