// one instantiation of the fused path kernel per translation unit (parallel builds)
// legacy_hybrid block bootstrap (bootstrap.py:310-458, engine.py:50-80), throughput build
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_bootstrap_all, float, LETF_PREC_F32, 32, 6, 4, false, letf::TM_ALL, 0, 0, 128, 1)
