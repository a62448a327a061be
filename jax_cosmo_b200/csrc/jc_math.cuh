// jc_math.cuh -- FP64 elementary functions for the hot kernels (sm_100a).
//
// Why not CUDA libm: ncu's source page for jc_power_kernel showed 731 warp-instructions per point of
// which only 290 were FP64 -- ~240 were UMOV / IMAD.MOV pairs materialising the 64-bit polynomial
// coefficients of exp/log/sin as immediates (profiles/r01_power_sass_mix.md).  Here every coefficient
// lives in __constant__ memory, so it is a c[bank][offset] operand of the DFMA itself, and the
// special-case paths (NaN/Inf/denormal/huge-argument) that this path never reaches are dropped.
// Accuracy (tests/test_gpu_parity.py::test_device_math): <= 2 ulp on the argument ranges documented
// per function; the project parity bar is rtol 1e-6 on C_ell.
#pragma once
#include <cuda_runtime.h>

struct JcMathK {
  double log2e, magic, ln2_hi, ln2_lo;
  double ec[13];  // 1/n!, n = 0..12
  double lg[7];   // fdlibm e_log.c Lg1..Lg7
  double invpio2, pio2_1, pio2_2, pio2_2t;
  double sc[12];  // [0..5] = S1..S6 (k_sin.c), [6..11] = C1..C6 (k_cos.c)
  double third, half, one, two, exp_lo;
};

static __constant__ JcMathK JCK = {
    1.4426950408889634074, 6755399441055744.0, 6.93147180369123816490e-01, 1.90821492927058770002e-10,
    {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
     1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600},
    {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
     1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01},
    6.36619772367581382433e-01, 1.57079632673412561417e+00, 6.07710050630396597660e-11,
    2.02226624879595063154e-21,
    {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
     2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
     4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
     -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11},
    1.0 / 3.0, 0.5, 1.0, 2.0, -708.0};

// 1/x for finite positive normal x: MUFU.RCP64H seed + 2 Newton steps, no special cases.
__device__ __forceinline__ double jcm_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, JCK.one);
  r = fma(r, e, r);
  e = fma(-x, r, JCK.one);
  r = fma(r, e, r);
  return r;
}

// x < -708 -> -708 as an integer compare on the high word (negative doubles order like unsigned integers):
// ISETP + 2 SEL instead of fmax's DSETP.MAX / FSEL / SEL / NaN fix-up / register moves.
__device__ __forceinline__ double jcm_clamp_exp_arg(double x) {
  const bool low = (unsigned)__double2hiint(x) > 0xC0862000u;  // hi word of -708.0
  return low ? JCK.exp_lo : x;
}

// exp(x) for x <= 709.  x < -708 is clamped (returns ~3e-308 instead of a denormal / 0).
__device__ __forceinline__ double jcm_exp(double x) {
  x = jcm_clamp_exp_arg(x);
  const double kd = fma(x, JCK.log2e, JCK.magic);
  const int k = __double2loint(kd);
  const double kf = kd - JCK.magic;
  double r = fma(kf, -JCK.ln2_hi, x);
  r = fma(kf, -JCK.ln2_lo, r);
  double p = JCK.ec[12];
#pragma unroll
  for (int i = 11; i >= 0; --i) p = fma(p, r, JCK.ec[i]);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// log(x) for finite positive normal x (fdlibm e_log.c without the special cases).
__device__ __forceinline__ double jcm_log(double x) {
  int hx = __double2hiint(x);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int i = (hx + 0x95f64) & 0x100000;  // mantissa > sqrt(2): halve, k += 1
  k += i >> 20;
  const double m = __hiloint2double(hx | (i ^ 0x3ff00000), __double2loint(x));
  const double f = m - JCK.one;
  const double s = f * jcm_rcp(JCK.two + f);
  const double z = s * s;
  double R = JCK.lg[6];
#pragma unroll
  for (int j = 5; j >= 0; --j) R = fma(R, z, JCK.lg[j]);
  R *= z;
  const double hfsq = JCK.half * f * f;
  const double dk = (double)k;
  // dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f)
  const double t = fma(s, hfsq + R, dk * JCK.ln2_lo);
  return fma(dk, JCK.ln2_hi, f - (hfsq - t));
}

// ---- table-driven variants (tables: plan.math_tab, staged in shared memory by the caller) -------------
// layout of the table: [0,256) 2^(j/256);  [256, 256+256) pairs {c_j, -ln c_j}, j = top 7 mantissa bits,
// c_j = 1/(1 + (j+1/2)/128) (c_0 = 1 so that log stays relatively accurate next to 1).
#define JCM_EXP_BITS 8
#define JCM_EXP_N (1 << JCM_EXP_BITS)
#define JCM_TAB_EXP 0
#define JCM_TAB_LOG JCM_EXP_N
#define JCM_TAB_DOUBLES (JCM_EXP_N + 256)

struct JcMathT {
  double k32, magic, l32_hi, l32_lo;  // N/ln2, 1.5*2^52, ln2/N split (N = JCM_EXP_N table entries per octave)
  double e[5];                        // 1/2, 1/6, 1/24, 1/120, 1/720
  double l[6];                        // -1/2, 1/3, -1/4, 1/5, -1/6, 1/7
};
static __constant__ JcMathT JCT = {
    1.4426950408889634074 * JCM_EXP_N, 6755399441055744.0, 6.93147180369123816490e-01 / JCM_EXP_N,
    1.90821492927058770002e-10 / JCM_EXP_N,
    {0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720},
    {-0.5, 1.0 / 3, -0.25, 0.2, -1.0 / 6, 1.0 / 7}};

// exp(x), x <= 709 (x < -708 clamped): 2^n * T[j] * p(r), 256 table entries per octave, |r| <= ln2/512,
// degree-4 Taylor (truncation r^5/120 < 3.8e-17; a 32-entry table needs degree 6: two more dependent DFMAs).
template <bool CLAMP = true>
__device__ __forceinline__ double jcm_exp_t(double x, const double* __restrict__ tab) {
  if (CLAMP) x = jcm_clamp_exp_arg(x);  // CLAMP = false: caller guarantees |x| < 700 (3 issue slots less)
  const double kd = fma(x, JCT.k32, JCT.magic);
  const int k = __double2loint(kd);
  const double kf = kd - JCT.magic;
  double r = fma(kf, -JCT.l32_hi, x);
  r = fma(kf, -JCT.l32_lo, r);
  double p = fma(JCT.e[2], r, JCT.e[1]);
  p = fma(p, r, JCT.e[0]);
  p = fma(p, r, JCK.one);
  p = fma(p, r, JCK.one);
  p *= tab[JCM_TAB_EXP + (k & (JCM_EXP_N - 1))];
  return __hiloint2double(__double2hiint(p) + ((k >> JCM_EXP_BITS) << 20), __double2loint(p));
}

// log(x) for finite normal x >= 1 (absolute error ~1e-16 max(1, |log x|); relative next to x = 1).
__device__ __forceinline__ double jcm_log_t(double x, const double* __restrict__ tab) {
  const int hx = __double2hiint(x);
  const int e = (hx >> 20) - 1023;
  const int j = (hx >> 13) & 127;
  const double m = __hiloint2double((hx & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 cl = *reinterpret_cast<const double2*>(tab + JCM_TAB_LOG + 2 * j);
  const double r = fma(m, cl.x, -JCK.one);
  double p = fma(JCT.l[5], r, JCT.l[4]);
  p = fma(p, r, JCT.l[3]);
  p = fma(p, r, JCT.l[2]);
  p = fma(p, r, JCT.l[1]);
  p = fma(p, r, JCT.l[0]);
  p = fma(p * r, r, r);  // r - r^2/2 + ...
  const double de = (double)e;
  return fma(de, JCK.ln2_hi, cl.y) + fma(de, JCK.ln2_lo, p);
}

// Reduced-degree variants for the tabulated power kernel (error budget 1e-10 per factor against a 1e-6 parity bar):
// exp: degree-3 Taylor on |r| <= ln2/512 (truncation r^4/24 < 1.4e-13) and a one-word ln2/N (|x| 1.1e-16 absolute);
// log: degree-4 series on |r| <= 1/256 (truncation r^5/5 < 1.9e-13) and a one-word ln2.
template <bool CLAMP = true>
__device__ __forceinline__ double jcm_exp_t3(double x, const double* __restrict__ tab) {
  if (CLAMP) x = jcm_clamp_exp_arg(x);
  const double kd = fma(x, JCT.k32, JCT.magic);
  const int k = __double2loint(kd);
  const double kf = kd - JCT.magic;
  const double r = fma(kf, -JCT.l32_hi, x);
  double p = fma(JCT.e[1], r, JCT.e[0]);
  p = fma(p, r, JCK.one);
  p = fma(p, r, JCK.one);
  p *= tab[JCM_TAB_EXP + (k & (JCM_EXP_N - 1))];
  return __hiloint2double(__double2hiint(p) + ((k >> JCM_EXP_BITS) << 20), __double2loint(p));
}
__device__ __forceinline__ double jcm_log_t4(double x, const double* __restrict__ tab) {
  const int hx = __double2hiint(x);
  const int e = (hx >> 20) - 1023;
  const int j = (hx >> 13) & 127;
  const double m = __hiloint2double((hx & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 cl = *reinterpret_cast<const double2*>(tab + JCM_TAB_LOG + 2 * j);
  const double r = fma(m, cl.x, -JCK.one);
  double p = fma(JCT.l[2], r, JCT.l[1]);
  p = fma(p, r, JCT.l[0]);
  p = fma(p * r, r, r);  // r - r^2/2 + r^3/3 - r^4/4
  return fma((double)e, JCK.ln2_hi, cl.y) + p;
}

// sin(x) for 0 <= x < ~1e6 (3-term Cody-Waite reduction; error grows linearly with x beyond 2^20*pi/2,
// where this path's integrand is already suppressed by > 1e-12).
__device__ __forceinline__ double jcm_sin(double x) {
  const double nd = fma(x, JCK.invpio2, JCK.magic);
  const int n = __double2loint(nd);
  const double nf = nd - JCK.magic;
  double r = fma(nf, -JCK.pio2_1, x);
  r = fma(nf, -JCK.pio2_2, r);
  r = fma(nf, -JCK.pio2_2t, r);
  const double z = r * r;
  const bool odd = n & 1;  // odd quadrant: cosine polynomial
  const double* c = JCK.sc + (odd ? 6 : 0);
  double p = c[5];
#pragma unroll
  for (int j = 4; j >= 0; --j) p = fma(p, z, c[j]);
  const double head = odd ? fma(z, -JCK.half, JCK.one) : r;
  const double mult = odd ? z * z : r * z;
  const double v = fma(mult, p, head);
  return (n & 2) ? -v : v;
}

// x^(-1/3) for positive x within float range: MUFU.LG2/EX2 seed (rel. err ~1e-6) + 2 Newton steps.
__device__ __forceinline__ double jcm_rcbrt(double x) {
  float lf, yf;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lf) : "f"((float)x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(-0.333333333f * lf));
  double y = (double)yf;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-x, y * y * y, JCK.one);  // 1 - x y^3
    y = fma(y * JCK.third, e, y);
  }
  return y;
}
