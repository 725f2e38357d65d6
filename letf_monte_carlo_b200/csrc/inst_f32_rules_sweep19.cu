// one instantiation of the fused path kernel per translation unit (parallel builds)
// rule kernel of the split strategy pipeline, the reference's default 19-strategy list on TQQQ / SSO / TMF / SPY: reads the
// producer's per-day records (MARKET = 5, letf_core.cuh kMarketFSeries), one trial per thread.
// 64-thread blocks x 5 per SM (10 warps, 204-register cap): measured 66 -> 63 ms per 227 328 trials against 128 x 2 (8 warps,
// 252 registers) and 65 ms for 128 x 3 (168 registers, 120 B of spills); the first registered build is the default
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_rules_sweep19, float, LETF_PREC_F32, 32, 6, 5, false, letf::slot_list_mask(2), 2, 29, 64, 5)
// A/B variants (LETF_KERNEL=<name>)
LETF_INSTANTIATE(f32_rules_sweep19_b128, float, LETF_PREC_F32, 32, 6, 5, false, letf::slot_list_mask(2), 2, 29, 128, 2)
LETF_INSTANTIATE(f32_rules_sweep19_r168, float, LETF_PREC_F32, 32, 6, 5, false, letf::slot_list_mask(2), 2, 29, 128, 3)
