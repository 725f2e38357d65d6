// one instantiation of the fused path kernel per translation unit (parallel builds)
// rule kernel of the split strategy pipeline, the reference's default 19-strategy list on TQQQ / SSO / TMF / SPY: reads the
// producer's per-day records (MARKET = 5, letf_core.cuh kMarketFSeries), one trial per thread
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_rules_sweep19, float, LETF_PREC_F32, 32, 6, 5, false, letf::slot_list_mask(2), 2, 29, 128, 2)
