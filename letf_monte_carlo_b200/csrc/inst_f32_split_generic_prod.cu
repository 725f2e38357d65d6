// one instantiation of the fused path kernel per translation unit (parallel builds)
// producer of the generic split strategy pipeline (letf_jfast.cuh EMIT, uncompacted slots): all six ETF series of
// cfg.ASSETS in their own slots, the ones outside the asset mask skipped at run time; TQQQ / UPRO / SSO / TMF t(5), SPY / QQQ t(10)
#include "letf_kernels.cuh"
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_all, 32, 6, 6, 0x30, true, 128, 2)
