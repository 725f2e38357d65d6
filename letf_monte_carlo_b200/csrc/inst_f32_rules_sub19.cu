// one instantiation of the fused path kernel per translation unit (parallel builds)
// rule kernel of the split strategy pipeline for any SUBSET of the reference's default 19-strategy list (slot list 3 =
// list 2 with routed slots: a slot can be switched off and writes the output row the request gives it)
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_rules_sub19, float, LETF_PREC_F32, 32, 6, 5, false, letf::slot_list_mask(3), 3, 29, 64, 5)
