// one instantiation of the fused path kernel per translation unit (parallel builds)
// legacy_hybrid block bootstrap, fp32 with tape / record / journal plumbing
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_bootstrap_io, float, LETF_PREC_F32, 32, 6, 4, true, letf::TM_ALL, 0, 0, 128, 1)
