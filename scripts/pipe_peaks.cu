// pipe_peaks.cu — issue-rate microbenchmark for the instruction classes the path kernels are made of (SURVEY.md 8d:
// "INT32 issue width on sm_100 must be measured by an IMAD/LOP3 microbenchmark, not assumed").
//
// Each thread runs ILP independent dependency chains of ONE instruction class for ITER iterations; every SM is filled to
// its thread limit, so the scheduler always has eligible warps and the measured rate is the pipe's issue limit.  Rates are
// reported in lane-operations per clock per SM (128 = one warp instruction per scheduler per clock), with the clock taken
// from clock64() inside the kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peaks pipe_peaks.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

constexpr int ILP = 8, ITER = 4096, BLK = 256;

enum Op { FFMA, FMUL, FMNMX, IMAD, IMAD_WIDE, LOP3, IADD3, MUFU_EX2, MUFU_RSQ, MUFU_LG2, MUFU_SIN, I2FP, F2F_64, DFMA, SHF };

template <int OP> __device__ __forceinline__ void step(uint32_t (&u)[ILP], float (&f)[ILP], double (&d)[ILP], uint32_t k) {
#pragma unroll
  for (int j = 0; j < ILP; j++) {
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(1.0000001f), "f"(1e-9f));
    if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(1.0000001f));
    if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(f[(j + 1) % ILP]));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[j]) : "r"(k), "r"(0x9E3779B9u));
    if (OP == IMAD_WIDE) {       // the Philox multiply: 32 x 32 -> 64, both halves consumed
      uint64_t p;
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(u[j]), "r"(0xD2511F53u));
      u[j] = (uint32_t)(p >> 32) ^ (uint32_t)p;
    }
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[j]) : "r"(k), "r"(u[(j + 1) % ILP]));
    if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[j]) : "r"(k));
    if (OP == MUFU_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
    if (OP == MUFU_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
    if (OP == MUFU_LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
    if (OP == MUFU_SIN) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
    if (OP == I2FP) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[j]) : "r"(u[j])); u[j] += __float_as_uint(f[j]); }
    if (OP == F2F_64) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[j]) : "f"(f[j])); f[j] += (float)__double2hiint(d[j]); }
    if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(1.0000001), "d"(1e-12));
    if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(u[j]) : "r"(u[(j + 1) % ILP]), "r"(k));
  }
}

__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

// cycles: per warp (t0, t1, smid).  The span of an SM is max(t1) - min(t0) over its warps: the schedulers are not fair
// (an old warp can run ahead greedily), so a single warp's own interval says nothing about the pipe.
template <int OP> __global__ void __launch_bounds__(BLK) bench(uint32_t* out, long long* cycles, uint32_t k) {
  uint32_t u[ILP]; float f[ILP]; double d[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) { u[j] = threadIdx.x * 2654435761u + j + k; f[j] = 1.0f + 1e-3f * (threadIdx.x + j); d[j] = 1.0 + 1e-3 * j; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < ITER; i++) step<OP>(u, f, d, k);
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) acc ^= u[j] ^ __float_as_uint(f[j]) ^ (uint32_t)__double2loint(d[j]);
  out[blockIdx.x * BLK + threadIdx.x] = acc;
  if ((threadIdx.x & 31) == 0) {
    long long* c = cycles + 3 * ((size_t)blockIdx.x * (BLK / 32) + (threadIdx.x >> 5));
    c[0] = t0; c[1] = t1; c[2] = smid();
  }
}

template <int OP> static double run(const char* name, int extra_per_op, int sms, int blocks_per_sm, uint32_t* out, long long* cyc, FILE* js, bool last) {
  const int grid = sms * blocks_per_sm;
  bench<OP><<<grid, BLK>>>(out, cyc, 3);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<grid, BLK>>>(out, cyc, 5);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const int nw = grid * (BLK / 32);
  long long* h = new long long[3 * (size_t)nw];
  cudaMemcpy(h, cyc, sizeof(long long) * 3 * nw, cudaMemcpyDeviceToHost);
  // per SM: warps it ran, first start, last end
  double rate_sum = 0; int sm_seen = 0;
  for (int sm = 0; sm < 4096; sm++) {
    long long lo = 0, hi = 0; int cnt = 0;
    for (int w = 0; w < nw; w++) if (h[3 * w + 2] == sm) { if (!cnt || h[3 * w] < lo) lo = h[3 * w]; if (!cnt || h[3 * w + 1] > hi) hi = h[3 * w + 1]; cnt++; }
    if (cnt) { rate_sum += (double)cnt * 32 * ILP * ITER / (double)(hi - lo); sm_seen++; }
  }
  delete[] h;
  const double mean = (double)blocks_per_sm * BLK * ILP * ITER / (rate_sum / sm_seen);
  const double lane_ops_per_clk_sm = rate_sum / sm_seen;
  printf("%-10s %8.2f lane-ops/clk/SM  (%5.3f warp-instr/clk/scheduler; %.3f ms, %.0f cycles per SM, +%d helper instr per op)\n", name,
         lane_ops_per_clk_sm, lane_ops_per_clk_sm / 128.0, ms, mean, extra_per_op);
  fprintf(js, "  \"%s\": {\"lane_ops_per_clk_per_sm\": %.2f, \"warp_instr_per_clk_per_scheduler\": %.4f, \"helper_instr_per_op\": %d}%s\n", name,
          lane_ops_per_clk_sm, lane_ops_per_clk_sm / 128.0, extra_per_op, last ? "" : ",");
  return lane_ops_per_clk_sm;
}

int main(int argc, char** argv) {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount, bps = 2048 / BLK;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, sizeof(uint32_t) * sms * bps * BLK); cudaMalloc(&cyc, sizeof(long long) * 3 * sms * bps * (BLK / 32));
  FILE* js = fopen(argc > 1 ? argv[1] : "pipe_peaks.json", "w");
  fprintf(js, "{\n  \"device\": \"%s\", \"sms\": %d, \"threads_per_sm\": %d, \"ilp\": %d,\n", pr.name, sms, bps * BLK, ILP);
  run<FFMA>("FFMA", 0, sms, bps, out, cyc, js, false);
  run<FMUL>("FMUL", 0, sms, bps, out, cyc, js, false);
  run<FMNMX>("FMNMX", 0, sms, bps, out, cyc, js, false);
  run<IMAD>("IMAD", 0, sms, bps, out, cyc, js, false);
  run<IMAD_WIDE>("IMAD_WIDE", 1, sms, bps, out, cyc, js, false);
  run<LOP3>("LOP3", 0, sms, bps, out, cyc, js, false);
  run<IADD3>("IADD3", 0, sms, bps, out, cyc, js, false);
  run<SHF>("SHF", 0, sms, bps, out, cyc, js, false);
  run<MUFU_EX2>("MUFU_EX2", 0, sms, bps, out, cyc, js, false);
  run<MUFU_RSQ>("MUFU_RSQ", 0, sms, bps, out, cyc, js, false);
  run<MUFU_LG2>("MUFU_LG2", 0, sms, bps, out, cyc, js, false);
  run<MUFU_SIN>("MUFU_SIN", 1, sms, bps, out, cyc, js, false);
  run<I2FP>("I2FP", 1, sms, bps, out, cyc, js, false);
  run<F2F_64>("F2F_F64_F32", 2, sms, bps, out, cyc, js, false);
  run<DFMA>("DFMA", 0, sms, bps, out, cyc, js, true);
  fprintf(js, "}\n");
  fclose(js);
  return 0;
}
