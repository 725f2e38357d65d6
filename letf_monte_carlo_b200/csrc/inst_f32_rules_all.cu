// one instantiation of the fused path kernel per translation unit (parallel builds)
// rule kernel of the generic split strategy pipeline: any strategy list over the six series (dynamic slots), reading the
// producer's per-day records (MARKET = 5, letf_core.cuh kMarketFSeries), one trial per thread
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_rules_all, float, LETF_PREC_F32, 32, 6, 5, false, letf::TM_ALL, 0, 0x3F, 128, 2)
