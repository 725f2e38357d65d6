// one instantiation of the fused path kernel per translation unit (parallel builds)
// antithetic-pair kernels of the lean institutional buy&hold shapes (letf_jfast.cuh)
#include "letf_kernels.cuh"
// BASELINE config 3: SPY + QQQ (+ TLT benchmark), both tracking errors t(10)
// (capped at 128 registers, no spills: four resident blocks per SM = 4 warps per scheduler, +12.7 % over the 142-register
//  three-block build, which stays selectable for A/B runs: LETF_KERNEL=f32_joint_t_pair_c3_r168)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_c3, 4, 3, 2, 3, false, 128, 4)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_c3_r168, 4, 3, 2, 3, false, 128, 3)
// S1 S2 S3 and the like: two LETFs t(5) + SPY t(10)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_3a, 4, 3, 3, 4, false, 128, 3)
// one series
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_1t5, 4, 3, 1, 0, false, 128, 3)
LETF_INSTANTIATE_JPAIR(f32_joint_t_pair_1t10, 4, 3, 1, 1, false, 128, 3)
