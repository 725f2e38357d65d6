// one instantiation of the fused path kernel per translation unit (parallel builds)
// rule kernel of the generic split strategy pipeline for lists of at most eight strategies: the slot loop is unrolled, so
// the slot state lives in registers instead of local memory (dynamic dispatch on p.strat[s].type per slot).
// 255 registers, 2 blocks per SM: a 168-register / 3-block build (252 B of spills) measured 3 % slower on the five-strategy list
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_rules_all8, float, LETF_PREC_F32, 8, 6, 5, false, letf::TM_ALL, 0, 0x3F, 128, 2)
