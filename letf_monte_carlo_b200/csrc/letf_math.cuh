// letf_math.cuh — precision-templated math + counter-based RNG for the fused LETF path kernel.
//
// Two arithmetic modes (template parameter R):
//   double : parity mode, reference arithmetic (the reference is float64 NumPy throughout)
//   float  : throughput mode, MUFU fast paths on the device
// Everything here is __host__ __device__ so the identical source also compiles with a plain
// host compiler for the test-only emulation build (tests/emu); the product only ever runs
// the device side.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define LETF_HD __host__ __device__ __forceinline__
#define LETF_HDN __host__ __device__ __noinline__
#else
#define LETF_HD inline
#define LETF_HDN
#endif

namespace letf {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (c0,c1,c2,c3), key = (k0,k1).
struct U4 { uint32_t x, y, z, w; };

LETF_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  // one IMAD.WIDE.U32 (a quarter-rate instruction, profiles/pipe_peaks.json).  Written as lo = a * b, hi = __umulhi(a, b) the
  // compiler split ~40 % of the Philox multiplies of the pair kernel's window loop into IMAD + IMAD.HI: the same pipe time,
  // two issue slots instead of one
  uint64_t p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#else
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

LETF_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0, l0, h1, l1;
    mulhilo(0xD2511F53u, c.x, h0, l0);
    mulhilo(0xCD9E8D57u, c.z, h1, l1);
    U4 n;
    n.x = h1 ^ c.y ^ k0;
    n.y = l1;
    n.z = h0 ^ c.w ^ k1;
    n.w = l0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// stream tags (counter word 3 = tag << 8 | sub-block)
enum : uint32_t {
  STREAM_START = 1, STREAM_SPELL = 2, STREAM_MKT0 = 3, STREAM_MKT1 = 4, STREAM_STRESS = 5,
  STREAM_JUMP = 6, STREAM_MKT_PAIR = 7 /* iid-normal market: one block per two days */,
  STREAM_TE0 = 8 /* +asset */, STREAM_PT0 = 16 /* +component, PARAM_T and the second noise blend of BOOTSTRAP */,
  STREAM_BOOT_A = 24 /* counter = block index: length, momentum A, block A, offset A */,
  STREAM_BOOT_B = 25 /* momentum B, block B, offset B */, STREAM_BN0 = 26 /* +component: sample_returns noise */,
  STREAM_BOOT_SYN = 30 /* +pool: synthetic blocks of an empty pool, counter = block index * block_size + row */
};

struct RngKey {
  uint32_t k0, k1;   // user seed
  uint32_t t0, t1;   // sim_id (trial id with the antithetic bit cleared)
};

LETF_HD U4 rng_block(const RngKey& k, uint32_t day, uint32_t stream, uint32_t sub = 0) {
  U4 c;
  c.x = k.t0; c.y = k.t1; c.z = day; c.w = (stream << 8) | sub;
  return philox4x32_10(c, k.k0, k.k1);
}

// N independent blocks of one (trial, day) with their rounds interleaved in program order: the 10 rounds of a
// block are a dependent IMAD.WIDE -> LOP3 chain (~10 cycles per round), N chains side by side keep the
// multiplier and logic pipes busy from a single warp.  Same values as N calls of rng_block.
template <int N>
LETF_HD void rng_blocks(const RngKey& k, uint32_t day, const uint32_t (&stream)[N], U4 (&out)[N]) {
  U4 c[N];
#pragma unroll
  for (int j = 0; j < N; j++) { c[j].x = k.t0; c[j].y = k.t1; c[j].z = day; c[j].w = stream[j] << 8; }
  uint32_t k0 = k.k0, k1 = k.k1;
#pragma unroll
  for (int r = 0; r < 10; r++) {
#pragma unroll
    for (int j = 0; j < N; j++) {
      uint32_t h0, l0, h1, l1;
      mulhilo(0xD2511F53u, c[j].x, h0, l0);
      mulhilo(0xCD9E8D57u, c[j].z, h1, l1);
      U4 n;
      n.x = h1 ^ c[j].y ^ k0;
      n.y = l1;
      n.z = h0 ^ c[j].w ^ k1;
      n.w = l0;
      c[j] = n;
    }
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int j = 0; j < N; j++) out[j] = c[j];
}

// ------------------------------------------------------------------------------------------
template <typename R> struct Mth;

template <> struct Mth<double> {
  typedef double T;
  static LETF_HD double sqrt_(double x) { return ::sqrt(x); }
  static LETF_HD double rsqrt_(double x) { return 1.0 / ::sqrt(x); }
  static LETF_HD double log_(double x) { return ::log(x); }
  static LETF_HD double neg2log_(double x) { return -2.0 * ::log(x); }
  static LETF_HD double tanh_(double x) { return ::tanh(x); }
  static LETF_HD double exp_(double x) { return ::exp(x); }
  static LETF_HD double floor_(double x) { return ::floor(x); }
  static LETF_HD double pow_(double x, double y) { return ::pow(x, y); }
  static LETF_HD double abs_(double x) { return ::fabs(x); }
  static LETF_HD double div_(double a, double b) { return a / b; }
  static LETF_HD double fma_(double a, double b, double c) { return ::fma(a, b, c); }
  static LETF_HD double u01(uint32_t a) { return ((double)a + 0.5) * 2.3283064365386963e-10; }
  static LETF_HD double u01_16(uint32_t a) { return ((double)a + 0.5) * (1.0 / 65536.0); }
  static LETF_HD double u01_12(uint32_t a) { return ((double)a + 0.5) * (1.0 / 4096.0); }
  static LETF_HD double bm_parts(uint32_t a, uint32_t b, double& cs, double& sn) {
    double th = 6.283185307179586476925 * u01(b);
    cs = ::cos(th); sn = ::sin(th);
    return -2.0 * ::log(u01(a));
  }
  static LETF_HD double bm_cos(uint32_t a, uint32_t b, double& cs) {
    cs = ::cos(6.283185307179586476925 * u01(b));
    return -2.0 * ::log(u01(a));
  }
  static LETF_HD void box_muller(uint32_t a, uint32_t b, double& n0, double& n1) {
    double r = ::sqrt(-2.0 * ::log(u01(a)));
    double th = 6.283185307179586476925 * u01(b);
    n0 = r * ::cos(th);
    n1 = r * ::sin(th);
  }
};

template <> struct Mth<float> {
  typedef float T;
#if defined(__CUDA_ARCH__)
  static LETF_HD float sqrt_(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static LETF_HD float rsqrt_(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static LETF_HD float log2_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static LETF_HD float log_(float x) { return 0.6931471805599453f * log2_(x); }
  static LETF_HD float neg2log_(float x) { return -1.3862943611198906f * log2_(x); }   // -2 ln x, one multiply
  static LETF_HD float tanh_(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static LETF_HD float pow_(float x, float y) { return __powf(x, y); }
  // a / b as one MUFU.RCP and one multiply (div.approx carries a range fix-up we do not need: |b| << 2^126 here)
  static LETF_HD float div_(float a, float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return a * r; }
#else
  static LETF_HD float sqrt_(float x) { return ::sqrtf(x); }
  static LETF_HD float rsqrt_(float x) { return 1.0f / ::sqrtf(x); }
  static LETF_HD float log2_(float x) { return ::log2f(x); }
  static LETF_HD float log_(float x) { return ::logf(x); }
  static LETF_HD float neg2log_(float x) { return -2.0f * ::logf(x); }
  static LETF_HD float tanh_(float x) { return ::tanhf(x); }
  static LETF_HD float pow_(float x, float y) { return ::powf(x, y); }
  static LETF_HD float div_(float a, float b) { return a / b; }
#endif
  static LETF_HD float exp_(float x) { return ::expf(x); }
  static LETF_HD float floor_(float x) { return ::floorf(x); }
  static LETF_HD float abs_(float x) { return ::fabsf(x); }
  static LETF_HD float fma_(float a, float b, float c) { return ::fmaf(a, b, c); }
  // (a + 0.5) * 2^-32, never 0; may round to exactly 1 (log -> 0, harmless)
  static LETF_HD float u01(uint32_t a) { return fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }
  static LETF_HD float u01_16(uint32_t a) { return fmaf((float)a, 1.0f / 65536.0f, 0.5f / 65536.0f); }
  static LETF_HD float u01_12(uint32_t a) { return fmaf((float)a, 1.0f / 4096.0f, 0.5f / 4096.0f); }
  // squared radius L = -2 ln U and the unit direction (cos, sin) of a Box-Muller pair
  static LETF_HD float bm_parts(uint32_t a, uint32_t b, float& cs, float& sn) {
    float th = (float)(int32_t)b * 1.4629180792671596e-9f;
#if defined(__CUDA_ARCH__)
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(th));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(th));
    return -1.3862943611198906f * log2_(u01(a));
#else
    cs = ::cosf(th); sn = ::sinf(th);
    return -2.0f * ::logf(u01(a));
#endif
  }
  // squared radius and cos only (the sine enters the t-variate as sin^2 = 1 - cos^2)
  static LETF_HD float bm_cos(uint32_t a, uint32_t b, float& cs) {
    float th = (float)(int32_t)b * 1.4629180792671596e-9f;
#if defined(__CUDA_ARCH__)
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(th));
    return -1.3862943611198906f * log2_(u01(a));
#else
    cs = ::cosf(th);
    return -2.0f * ::logf(u01(a));
#endif
  }
  static LETF_HD void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    float u = u01(a);
    // angle in [-pi, pi): signed reinterpretation of the second word
    float th = (float)(int32_t)b * 1.4629180792671596e-9f;  // 2*pi / 2^32
#if defined(__CUDA_ARCH__)
    float r = sqrt_(-1.3862943611198906f * log2_(u));        // sqrt(-2 ln u) = sqrt(-2 ln2 * log2 u)
    float cs, sn;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(th));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(th));
    n0 = r * cs;
    n1 = r * sn;
#else
    float r = ::sqrtf(-2.0f * ::logf(u));
    n0 = r * ::cosf(th);
    n1 = r * ::sinf(th);
#endif
  }
};

LETF_HD double rmax(double a, double b) { return a > b ? a : b; }   // python max(a, b) for non-NaN b
LETF_HD double rmin(double a, double b) { return a < b ? a : b; }
LETF_HD float rmax(float a, float b) { return fmaxf(a, b); }        // same result whenever b is not NaN (it never is here)
LETF_HD float rmin(float a, float b) { return fminf(a, b); }
template <typename R> LETF_HD R clipr(R x, R lo, R hi) { return x < lo ? lo : (x > hi ? hi : x); } // NaN passes through
// clip for values that are known not to be NaN-significant (fp32 hot path): two FMNMX
LETF_HD float clipfast(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
LETF_HD double clipfast(double x, double lo, double hi) { return clipr(x, lo, hi); }
// clip of quantities that are finite by construction: exact NaN-preserving form in fp64, two FMNMX in fp32
LETF_HD float clipx(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
LETF_HD double clipx(double x, double lo, double hi) { return clipr(x, lo, hi); }
template <typename R> LETF_HD bool isnan_(R x) { return x != x; }
template <typename R> LETF_HD bool isfinite_(R x) { return (x - x) == (R)0; }

// ------------------------------------------------------------------------------------------
// chi-square(nu) and Student-t(df) samplers.
//
// Fast path (2*shape integral, nu an integer <= 11): Gamma(k/2) = -ln(prod of floor(k/2) uniforms)
// (+ N^2/2 when k is odd), i.e. exact in distribution and branch-free:
//     chi2(k) = -2 ln(U_1 ... U_m) + [k odd] N^2,  m = k / 2
// The m uniforms are cut from two 32-bit words: 32 bits each for m <= 2, 16 bits for m <= 4,
// 12 bits for m == 5.  General nu (calibrated joint models, calibration.py:399, give nu in
// [3.2, 12]): Marsaglia-Tsang (2000) squeeze/rejection with extra Philox blocks per retry.
template <typename R>
LETF_HD R chi2_int_fast(int k, uint32_t wa, uint32_t wb, R nrm) {
  typedef Mth<R> M;
  if (k == 5) {                        // cfg.ASSETS default for the LETFs: one uniform compare on the hot path
    R g5 = (R)-2 * M::log_(M::u01(wa) * M::u01(wb));
    return g5 + nrm * nrm;
  }
  int m = k >> 1;
  R prod = (R)1;
  if (m == 1) prod = M::u01(wa);
  else if (m == 2) prod = M::u01(wa) * M::u01(wb);
  else if (m == 3) prod = M::u01_16(wa >> 16) * M::u01_16(wa & 0xFFFFu) * M::u01_16(wb >> 16);
  else if (m == 4) prod = M::u01_16(wa >> 16) * M::u01_16(wa & 0xFFFFu) * M::u01_16(wb >> 16) * M::u01_16(wb & 0xFFFFu);
  else if (m == 5) prod = M::u01_12(wa >> 20) * M::u01_12((wa >> 8) & 0xFFFu) * M::u01_12(wb >> 20) *
                          M::u01_12((wb >> 8) & 0xFFFu) * M::u01_12(((wa & 0xFFu) << 4) | (wb & 0xFu));
  R g = m > 0 ? (R)-2 * M::log_(prod) : (R)0;
  if (k & 1) g += nrm * nrm;
  return g;
}

// Marsaglia-Tsang gamma(shape a >= 1) * 2 = chi2(2a).  First attempt uses (n0, u0) supplied by the
// caller's block; retries pull blocks (stream, sub = 1, 2, ...).
// Two calling conventions for the (rare, non-inlined) retry loop.  By value keeps the caller's Trial object out of
// memory: with a reference to its `key` member the object escapes and, in the large typed / generic builds, every
// parameter read of the day loop degrades from a constant-bank load to a generic load through the saved reference.
// By reference costs the lean 128-register builds nothing extra (their key then lives on the stack).
// one Marsaglia-Tsang attempt: squeeze test, then the log test
template <typename R>
LETF_HD bool mt_accept(R d, R c, R x, uint32_t uw, R& out) {
  typedef Mth<R> M;
  R v = (R)1 + c * x;
  if (v > (R)0) {
    v = v * v * v;
    R u = M::u01(uw);
    R x2 = x * x;
    if (u < (R)1 - (R)0.0331 * x2 * x2 || M::log_(u) < (R)0.5 * x2 + d * ((R)1 - v + M::log_(v))) { out = (R)2 * d * v; return true; }
  }
  return false;
}
// the retries (acceptance > 95 % per attempt, so ~4 % of the calls get here): blocks (stream, sub = 1, 2, ...)
template <typename R>
LETF_HD R chi2_mt_retry_body(R d, R c, const RngKey& key, uint32_t day, uint32_t stream) {
  typedef Mth<R> M;
  for (uint32_t sub = 1; sub < 64; sub++) {
    U4 b = rng_block(key, day, stream, sub);
    R x, dump, out;
    M::box_muller(b.x, b.y, x, dump);
    if (mt_accept<R>(d, c, x, b.z, out)) return out;
  }
  return (R)2 * d;  // unreachable in practice
}
template <typename R>
LETF_HDN R chi2_mt_retry_byval(R d, R c, const RngKey key, uint32_t day, uint32_t stream) {
  return chi2_mt_retry_body<R>(d, c, key, day, stream);
}
// the lean 128-register builds keep the whole sampler out of line, first attempt included (one loop, as compact as it gets:
// inlining the first attempt, or even restructuring this function, costs them hundreds of bytes of spills)
template <typename R>
LETF_HDN R chi2_mt_full_byref(R nu, R n0, uint32_t u0w, const RngKey& key, uint32_t day, uint32_t stream) {
  typedef Mth<R> M;
  R a = (R)0.5 * nu;
  R d = a - (R)(1.0 / 3.0);
  R c = M::rsqrt_((R)9 * d);
  R x = n0;
  uint32_t uw = u0w;
  for (uint32_t sub = 1; sub < 64; sub++) {
    R v = (R)1 + c * x;
    if (v > (R)0) {
      v = v * v * v;
      R u = M::u01(uw);
      R x2 = x * x;
      if (u < (R)1 - (R)0.0331 * x2 * x2) return (R)2 * d * v;
      if (M::log_(u) < (R)0.5 * x2 + d * ((R)1 - v + M::log_(v))) return (R)2 * d * v;
    }
    U4 b = rng_block(key, day, stream, sub);
    R dump;
    M::box_muller(b.x, b.y, x, dump);
    uw = b.z;
  }
  return (R)2 * d;  // unreachable in practice (acceptance > 95 % per attempt)
}
// the 6-asset builds run the first attempt inline (calibrated joint models have non-integer dof, calibration.py:399: this
// runs every generated day) and only a rejection pays for the call
template <typename R, bool KEYREF = false>
LETF_HD R chi2_marsaglia_tsang(R nu, R n0, uint32_t u0w, const RngKey& key, uint32_t day, uint32_t stream) {
  typedef Mth<R> M;
  if (KEYREF) return chi2_mt_full_byref<R>(nu, n0, u0w, key, day, stream);
  const R d = (R)0.5 * nu - (R)(1.0 / 3.0);
  const R c = M::rsqrt_((R)9 * d);
  R out;
  if (mt_accept<R>(d, c, n0, u0w, out)) return out;
  return chi2_mt_retry_byval<R>(d, c, key, day, stream);
}

struct DfInfo {      // host-precomputed per distribution
  int fast_k;        // integer dof in [1, 11] -> fast path, else 0
};

LETF_HD int df_fast_k(double df) {
  double r = ::floor(df + 0.5);
  return (::fabs(df - r) < 1e-12 && r >= 1.0 && r <= 11.0) ? (int)r : 0;
}

// Student-t(5) from one Philox block, straight-line (the default tracking-error dof of the LETFs):
// N0 * sqrt(5) * rsqrt(-2 ln(U1 U2) + N1^2).  Same arithmetic as student_t_block with fast_k == 5.
template <typename R>
LETF_HD R student_t5(const U4& b) {
  typedef Mth<R> M;
  if (sizeof(R) == 8) {
    R n0, n1;
    M::box_muller(b.x, b.y, n0, n1);
    R c2 = (R)-2 * M::log_(M::u01(b.z) * M::u01(b.w)) + n1 * n1;
    return n0 * M::sqrt_(M::div_((R)5, rmax(c2, (R)1e-30)));
  }
  // fp32: with L = -2 ln U (the squared Box-Muller radius), N0 = sqrt(L) cos, N1^2 = L sin^2, so
  //   t = cos * sqrt(5 L / c2) = cos * (5 L) * rsqrt(5 L * c2),   sin^2 = 1 - cos^2
  // -> 4 MUFU (lg2, cos, lg2, rsq)
  R cs;
  R L = M::bm_cos(b.x, b.y, cs);
  R c2 = M::fma_(L, M::fma_(-cs, cs, (R)1), M::neg2log_(M::u01(b.z) * M::u01(b.w)));
  R x5 = (R)5 * L;
  return cs * (x5 * M::rsqrt_(rmax(x5 * c2, (R)1e-30)));
}

// Student-t(10), straight-line (cfg.ASSETS default for the unlevered SPY / QQQ series): chi2(10) = -2 ln of
// five uniforms (12 bits each, cut from two words exactly as chi2_int_fast does), no N1^2 term:
//   t = cos * sqrt(10 L / c2) = cos * (10 L) * rsqrt(10 L * c2)
template <typename R>
LETF_HD R student_t10(const U4& b) {
  typedef Mth<R> M;
  const uint32_t wa = b.z, wb = b.w;
  R prod = M::u01_12(wa >> 20) * M::u01_12((wa >> 8) & 0xFFFu) * M::u01_12(wb >> 20) *
           M::u01_12((wb >> 8) & 0xFFFu) * M::u01_12(((wa & 0xFFu) << 4) | (wb & 0xFu));
  if (sizeof(R) == 8) {
    R n0, n1;
    M::box_muller(b.x, b.y, n0, n1);
    return n0 * M::sqrt_(M::div_((R)10, rmax((R)-2 * M::log_(prod), (R)1e-30)));
  }
  R cs;
  R L = M::bm_cos(b.x, b.y, cs);
  R x10 = (R)10 * L;
  return cs * (x10 * M::rsqrt_(rmax(x10 * M::neg2log_(prod), (R)1e-30)));
}

// Student-t(df) from one Philox block: N0 / sqrt(chi2(df) / df).
template <typename R, bool KEYREF = false>
LETF_HD R student_t_block(const U4& b, R df, int fast_k, const RngKey& key, uint32_t day, uint32_t stream) {
  typedef Mth<R> M;
  if (sizeof(R) == 4 && fast_k == 5) return student_t5<R>(b);
  if (sizeof(R) == 4 && fast_k == 10) return student_t10<R>(b);
  R n0, n1;
  M::box_muller(b.x, b.y, n0, n1);
  R c2 = fast_k ? chi2_int_fast<R>(fast_k, b.z, b.w, n1)
                : chi2_marsaglia_tsang<R, KEYREF>(df, n1, b.z, key, day, stream);
  if (sizeof(R) == 8) return n0 * M::sqrt_(M::div_(df, rmax(c2, (R)1e-30)));
  return n0 * (M::sqrt_(df) * M::rsqrt_(rmax(c2, (R)1e-30)));   // sqrt(df) is loop-invariant: one MUFU per variate
}

}  // namespace letf
