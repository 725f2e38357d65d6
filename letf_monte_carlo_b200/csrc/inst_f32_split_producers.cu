// one instantiation of the fused path kernel per translation unit (parallel builds)
// producers of the split strategy pipeline (letf_jfast.cuh EMIT): market + layers of one antithetic pair per thread,
// buy&hold statistics, and the per-day records the rule kernels consume
#include "letf_kernels.cuh"
// BASELINE config 4: TQQQ / TMF t(5), SPY t(10)
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_c4, 32, 4, 3, 4, true, 128, 3)
// the reference's default sweep: TQQQ / SSO / TMF t(5), SPY t(10)
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_sweep19, 32, 4, 4, 8, true, 128, 3)
// A/B variants capped at 128 registers / four blocks per SM (LETF_KERNEL=<name>)
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_c4_r128, 32, 4, 3, 4, true, 128, 4)
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_sweep19_r128, 32, 4, 4, 8, true, 128, 4)
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_c4_b64, 32, 4, 3, 4, true, 64, 7)
// calibrated tracking_residual_model (per-series dof, tanh cap): the tracking-error sampler dispatches at run time
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_c4_tedf, 32, 4, 3, 0x80, true, 128, 3)
LETF_INSTANTIATE_JPAIR(f32_joint_t_prod_sweep19_tedf, 32, 4, 4, 0x80, true, 128, 3)
