// one instantiation of the fused path kernel per translation unit (parallel builds)
// antithetic-pair kernels of the lean institutional buy&hold shapes (letf_jfast.cuh)
#include "letf_kernels.cuh"
// BASELINE config 3: SPY + QQQ (+ TLT benchmark), both tracking errors t(10)
// (142 registers, three resident blocks per SM.  Capped at 128 registers / four blocks it also compiles without spills, but
//  measures 5 % slower at the literal 1M-trial size, 164.0 vs 156.4 ms on the same box - A/B: LETF_KERNEL=f32_joint_t_pair_c3_r128;
//  at trial counts that quantise badly into three-block waves the capped build wins, see profiles/README.md)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_c3, 4, 3, 2, 3, false, 128, 3)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_c3_r128, 4, 3, 2, 3, false, 128, 4)
// 64-thread blocks, seven per SM: 14 warps instead of 12 at the same register budget (146 cap)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_c3_b64, 4, 3, 2, 3, false, 64, 7)
// S1 S2 S3 and the like: two LETFs t(5) + SPY t(10)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_3a, 4, 3, 3, 4, false, 128, 3)
// one series
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_1t5, 4, 3, 1, 0, false, 128, 3)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_1t10, 4, 3, 1, 1, false, 128, 3)
// calibrated tracking_residual_model (per-series dof, tanh cap), two / three series
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_2tedf, 4, 3, 2, 0x80, false, 128, 3)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_3tedf, 4, 3, 3, 0x80, false, 128, 3)
