// one instantiation of the fused path kernel per translation unit (parallel builds)
// materialised-path mode (run_strategy_fixed seam, strategy.py:55) with the reference's default 19-strategy list
// compiled in as typed slots: the kernel that streams caller-supplied series at HBM-relevant rates
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_series_sweep19, float, LETF_PREC_F32, 32, 6, 3, false, letf::slot_list_mask(2), 2, 0, 128, 1)
