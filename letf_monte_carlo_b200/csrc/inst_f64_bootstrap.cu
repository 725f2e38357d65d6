// one instantiation of the fused path kernel per translation unit (parallel builds)
// legacy_hybrid block bootstrap (bootstrap.py:310-458, engine.py:50-80), parity build
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f64_bootstrap, double, LETF_PREC_F64, 32, 6, 4, true, letf::TM_ALL, 0, 0, 128, 1)
