// jc_dual.cuh -- forward-mode (tangent) arithmetic for the JVP variant of the pipeline.
//
// Every kernel of the pipeline is a template on its scalar type T: `double` (the hot path) or `DualN<K>`
// (value + K directional derivatives; `Dual` = DualN<1>).  A dual pass yields exactly what jax.jacfwd gives for the
// reference's discretised program: interpolation brackets are fixed grids (no derivative), the
// halofit root index and every clip / abs / max decision is taken on the value and the derivative
// follows the selected branch ("frozen-index derivative").
//
// DualN<K> with K > 1 carries K tangents through ONE evaluation of the value: every exp / log / sin / rcp / sqrt of
// K1..K3 is computed once and each tangent costs one or two multiply-adds per operation (the 7-parameter Jacobian of
// BASELINE config 4 runs as groups of 4 + 3 directions instead of 7 passes that each recompute the value).  The tangent
// expressions are written per direction exactly as in the K = 1 case, so a grouped pass reproduces the one-direction
// pass.
//
// Workspace layout for dual passes: the value of element i of a table lives at p[i] (same plane the
// double pipeline uses) and tangent k at p[i + (k + 1) * doff] (further planes); `doff` is 0 for double.
#pragma once
#include "jc_math.cuh"

template <int K>
struct DualN {
  double v, d[K];
  __host__ __device__ DualN() {}
  __host__ __device__ DualN(double v_) : v(v_) {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = 0.0;
  }
};
using Dual = DualN<1>;

template <class T> struct JxTangents { static constexpr int N = 0; };
template <int K> struct JxTangents<DualN<K>> { static constexpr int N = K; };

#define JCD template <int K> __host__ __device__ __forceinline__
#define JCD_EACH _Pragma("unroll") for (int k = 0; k < K; ++k)
JCD DualN<K> operator+(DualN<K> a, DualN<K> b) { DualN<K> r; r.v = a.v + b.v; JCD_EACH r.d[k] = a.d[k] + b.d[k]; return r; }
JCD DualN<K> operator+(DualN<K> a, double b) { a.v = a.v + b; return a; }
JCD DualN<K> operator+(double a, DualN<K> b) { b.v = a + b.v; return b; }
JCD DualN<K> operator-(DualN<K> a, DualN<K> b) { DualN<K> r; r.v = a.v - b.v; JCD_EACH r.d[k] = a.d[k] - b.d[k]; return r; }
JCD DualN<K> operator-(DualN<K> a, double b) { a.v = a.v - b; return a; }
JCD DualN<K> operator-(double a, DualN<K> b) { DualN<K> r; r.v = a - b.v; JCD_EACH r.d[k] = -b.d[k]; return r; }
JCD DualN<K> operator-(DualN<K> a) { DualN<K> r; r.v = -a.v; JCD_EACH r.d[k] = -a.d[k]; return r; }
JCD DualN<K> operator*(DualN<K> a, DualN<K> b) {
  DualN<K> r;
  r.v = a.v * b.v;
  JCD_EACH r.d[k] = fma(a.v, b.d[k], a.d[k] * b.v);
  return r;
}
JCD DualN<K> operator*(DualN<K> a, double b) { DualN<K> r; r.v = a.v * b; JCD_EACH r.d[k] = a.d[k] * b; return r; }
JCD DualN<K> operator*(double a, DualN<K> b) { DualN<K> r; r.v = a * b.v; JCD_EACH r.d[k] = a * b.d[k]; return r; }
JCD DualN<K>& operator+=(DualN<K>& a, DualN<K> b) { a.v += b.v; JCD_EACH a.d[k] += b.d[k]; return a; }
JCD DualN<K>& operator*=(DualN<K>& a, DualN<K> b) { a = a * b; return a; }
JCD DualN<K>& operator*=(DualN<K>& a, double b) { a.v *= b; JCD_EACH a.d[k] *= b; return a; }
#undef JCD
#define JCD template <int K> __device__ __forceinline__

// value accessor that also accepts plain doubles; tangent seed (direction k of a parameter = row k of the tangent block)
__device__ __forceinline__ double jx_val(double a) { return a; }
JCD double jx_val(DualN<K> a) { return a.v; }
__device__ __forceinline__ void jx_seed(double&, const double*, int, int) {}
JCD void jx_seed(DualN<K>& x, const double* __restrict__ tangent, int i, int row_stride) {
  JCD_EACH x.d[k] = tangent[(size_t)k * row_stride + i];
}

// tangent k of `flag` := 1 when direction k has a component along a parameter the tracer kernels depend on (columns Omega_c,
// Omega_b, Omega_k, w0, wa, gamma of a cosmology row), else 0
__device__ __forceinline__ void jx_flag_moves_r(double&, const double*, int) {}
JCD void jx_flag_moves_r(DualN<K>& flag, const double* __restrict__ tangent, int ncp) {
  JCD_EACH {
    const double* t = tangent + (size_t)k * ncp;
    bool any = t[0] != 0.0 || t[1] != 0.0 || t[5] != 0.0 || t[6] != 0.0 || t[7] != 0.0;
    if (ncp > 8) any = any || t[8] != 0.0;
    flag.d[k] = any ? 1.0 : 0.0;
  }
}

// ---- elementary functions, overloaded for double and DualN ---------------------------------------------
__device__ __forceinline__ double jx_rcp(double x) { return jcm_rcp(x); }
JCD DualN<K> jx_rcp(DualN<K> x) {
  const double r = jcm_rcp(x.v);
  const double g = -(r * r);  // the factors shared by all tangents are formed once (here and below)
  DualN<K> o;
  o.v = r;
  JCD_EACH o.d[k] = g * x.d[k];
  return o;
}
JCD DualN<K> operator/(DualN<K> a, DualN<K> b) { return a * jx_rcp(b); }
JCD DualN<K> operator/(DualN<K> a, double b) { const double r = 1.0 / b; DualN<K> o; o.v = a.v * r; JCD_EACH o.d[k] = a.d[k] * r; return o; }
JCD DualN<K> operator/(double a, DualN<K> b) { return a * jx_rcp(b); }

__device__ __forceinline__ double jx_exp(double x) { return jcm_exp(x); }
JCD DualN<K> jx_exp(DualN<K> x) { const double e = jcm_exp(x.v); DualN<K> o; o.v = e; JCD_EACH o.d[k] = e * x.d[k]; return o; }
__device__ __forceinline__ double jx_log(double x) { return jcm_log(x); }
JCD DualN<K> jx_log(DualN<K> x) {
  DualN<K> o;
  o.v = jcm_log(x.v);
  const double r = jcm_rcp(x.v);
  JCD_EACH o.d[k] = x.d[k] * r;
  return o;
}
__device__ __forceinline__ double jx_exp_t(double x, const double* tab) { return jcm_exp_t(x, tab); }
JCD DualN<K> jx_exp_t(DualN<K> x, const double* tab) { const double e = jcm_exp_t(x.v, tab); DualN<K> o; o.v = e; JCD_EACH o.d[k] = e * x.d[k]; return o; }
// exp of a bounded argument (|x| < 700): no underflow clamp
__device__ __forceinline__ double jx_exp_tb(double x, const double* tab) { return jcm_exp_t<false>(x, tab); }
JCD DualN<K> jx_exp_tb(DualN<K> x, const double* tab) { const double e = jcm_exp_t<false>(x.v, tab); DualN<K> o; o.v = e; JCD_EACH o.d[k] = e * x.d[k]; return o; }
__device__ __forceinline__ double jx_log_t(double x, const double* tab) { return jcm_log_t(x, tab); }
JCD DualN<K> jx_log_t(DualN<K> x, const double* tab) {
  DualN<K> o;
  o.v = jcm_log_t(x.v, tab);
  const double r = jcm_rcp(x.v);
  JCD_EACH o.d[k] = x.d[k] * r;
  return o;
}
__device__ __forceinline__ double jx_sqrt(double x) { return sqrt(x); }
JCD DualN<K> jx_sqrt(DualN<K> x) { const double s = sqrt(x.v), g = 0.5 / s; DualN<K> o; o.v = s; JCD_EACH o.d[k] = g * x.d[k]; return o; }
__device__ __forceinline__ double jx_sin(double x) { return jcm_sin(x); }
JCD DualN<K> jx_sin(DualN<K> x) {  // cos(x) = sin(x + pi/2); arguments are >= 0 on this path
  DualN<K> o;
  o.v = jcm_sin(x.v);
  const double c = jcm_sin(x.v + 1.5707963267948966);
  JCD_EACH o.d[k] = c * x.d[k];
  return o;
}
__device__ __forceinline__ double jx_rcbrt(double x) { return jcm_rcbrt(x); }
JCD DualN<K> jx_rcbrt(DualN<K> x) {  // d x^(-1/3) = -1/3 x^(-4/3) = -1/3 y^4
  const double y = jcm_rcbrt(x.v);
  const double y2 = y * y, g = (-1.0 / 3.0) * y2 * y2;
  DualN<K> o;
  o.v = y;
  JCD_EACH o.d[k] = g * x.d[k];
  return o;
}
// x^y = exp(y ln x) for a positive normal base, with this file's exp / log (error ~ |y ln x| ulp, a few 1e-16 for the
// Eisenstein-Hu constants).  libm's pow costs ~300 instructions a call, and the ~20 calls of the EH-constant thread sat in
// front of the setup kernel's first barrier: 14 % of its warp time (ncu source page, BSSY at the barrier).
__device__ __forceinline__ double jx_pow(double x, double y) { return jcm_exp(y * jcm_log(x)); }
JCD DualN<K> jx_pow(DualN<K> x, double y) {
  const double p = jcm_exp(y * jcm_log(x.v)), g = y * p / x.v;
  DualN<K> o;
  o.v = p;
  JCD_EACH o.d[k] = g * x.d[k];
  return o;
}
JCD DualN<K> jx_pow(DualN<K> x, DualN<K> y) {
  const double lx = jcm_log(x.v);
  const double p = jcm_exp(y.v * lx), g = y.v / x.v;
  DualN<K> o;
  o.v = p;
  JCD_EACH o.d[k] = p * fma(y.d[k], lx, g * x.d[k]);
  return o;
}
// selections follow the value (the sub-gradient jax takes away from ties)
__device__ __forceinline__ double jx_max(double a, double b) { return fmax(a, b); }
JCD DualN<K> jx_max(DualN<K> a, double b) { return a.v >= b ? a : DualN<K>(b); }
JCD DualN<K> jx_max(DualN<K> a, DualN<K> b) { return a.v >= b.v ? a : b; }
__device__ __forceinline__ double jx_min(double a, double b) { return fmin(a, b); }
JCD DualN<K> jx_min(DualN<K> a, double b) { return a.v <= b ? a : DualN<K>(b); }
// max(x, 0) and, for x >= 0, max(x, 1) as integer selects on the words of x: 3 ALU instructions instead of
// fmax's DSETP.MAX / FSEL / SEL / NaN-fixup / register-move sequence (8-10 issue slots, one on the FP64
// pipe); the clips of the lensing-efficiency integrand (probes.py:49, background.py:242) run 3 per point.
// NaN passes through, -0 becomes +0.
__device__ __forceinline__ double jx_clip0(double x) {
  const int hi = __double2hiint(x), keep = ~(hi >> 31);
  return __hiloint2double(hi & keep, __double2loint(x) & keep);
}
JCD DualN<K> jx_clip0(DualN<K> a) { return a.v >= 0.0 ? a : DualN<K>(0.0); }
__device__ __forceinline__ double jx_floor1(double x) {  // x >= 0
  const int hi = __double2hiint(x);
  const bool small = hi < 0x3ff00000;
  return __hiloint2double(small ? 0x3ff00000 : hi, small ? 0 : __double2loint(x));
}
JCD DualN<K> jx_floor1(DualN<K> a) { return a.v >= 1.0 ? a : DualN<K>(1.0); }
__device__ __forceinline__ double jx_abs(double a) { return fabs(a); }
JCD DualN<K> jx_abs(DualN<K> a) { return a.v >= 0.0 ? a : -a; }
__device__ __forceinline__ double jx_fma(double a, double b, double c) { return fma(a, b, c); }
JCD DualN<K> jx_fma(DualN<K> a, DualN<K> b, DualN<K> c) { return a * b + c; }
JCD DualN<K> jx_fma(double a, DualN<K> b, DualN<K> c) { return a * b + c; }
JCD DualN<K> jx_fma(DualN<K> a, double b, DualN<K> c) { return a * b + c; }
JCD DualN<K> jx_fma(DualN<K> a, DualN<K> b, double c) { return a * b + c; }
JCD DualN<K> jx_fma(double a, DualN<K> b, double c) { return a * b + c; }
JCD DualN<K> jx_fma(DualN<K> a, double b, double c) { return a * b + c; }
JCD DualN<K> jx_fma(double a, double b, DualN<K> c) { c.v = fma(a, b, c.v); return c; }

__device__ __forceinline__ double jx_shfl_xor(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
JCD DualN<K> jx_shfl_xor(DualN<K> v, int o) {
  DualN<K> r;
  r.v = __shfl_xor_sync(0xffffffffu, v.v, o);
  JCD_EACH r.d[k] = __shfl_xor_sync(0xffffffffu, v.d[k], o);
  return r;
}

// ---- workspace access: value plane at p[i], tangent plane k at p[i + (k + 1) * doff] ----------------------
template <class T> struct JxMem;
template <> struct JxMem<double> {
  static __device__ __forceinline__ double ld(const double* p, ptrdiff_t) { return *p; }
  static __device__ __forceinline__ void st(double* p, ptrdiff_t, double x) { *p = x; }
};
template <int K> struct JxMem<DualN<K>> {
  static __device__ __forceinline__ DualN<K> ld(const double* p, ptrdiff_t doff) {
    DualN<K> r;
    r.v = p[0];
    JCD_EACH r.d[k] = p[(ptrdiff_t)(k + 1) * doff];
    return r;
  }
  static __device__ __forceinline__ void st(double* p, ptrdiff_t doff, DualN<K> x) {
    p[0] = x.v;
    JCD_EACH p[(ptrdiff_t)(k + 1) * doff] = x.d[k];
  }
};
#undef JCD
#undef JCD_EACH
