// one instantiation of the fused path kernel per translation unit (parallel builds)
// rule kernel of the split strategy pipeline, BASELINE config 4 list (S1 S2 S4 S5 S6 S19 on TQQQ / TMF / SPY): reads the
// producer's per-day records (MARKET = 5, letf_core.cuh kMarketFSeries), one trial per thread
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_rules_c4, float, LETF_PREC_F32, 8, 6, 5, false, letf::slot_list_mask(1), 1, 25, 128, 3)
