// one instantiation of the fused path kernel per translation unit (parallel builds)
// legacy_hybrid block bootstrap, lean buy&hold build (compacted asset slots, no strategy slots' work)
#include "letf_kernels.cuh"
LETF_INSTANTIATE(f32_bootstrap_bench, float, LETF_PREC_F32, 4, 3, 4, false, letf::TM_BENCH, 0, 0, 64, 6)
