// jc_dual.cuh -- forward-mode (tangent) arithmetic for the JVP variant of the pipeline.
//
// Every kernel of the pipeline is a template on its scalar type T: `double` (the hot path) or `Dual`
// (value + one directional derivative).  A Dual pass yields exactly what jax.jacfwd gives for the
// reference's discretised program: interpolation brackets are fixed grids (no derivative), the
// halofit root index and every clip / abs / max decision is taken on the value and the derivative
// follows the selected branch ("frozen-index derivative").
//
// Workspace layout for Dual passes: the value of element i of a table lives at p[i] (same plane the
// double pipeline uses) and its tangent at p[i + doff] (a second plane); `doff` is 0 for double.
#pragma once
#include "jc_math.cuh"

struct Dual {
  double v, d;
  __host__ __device__ Dual() {}
  __host__ __device__ Dual(double v_) : v(v_), d(0.0) {}
  __host__ __device__ Dual(double v_, double d_) : v(v_), d(d_) {}
};

#define JCD __host__ __device__ __forceinline__
JCD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
JCD Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
JCD Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
JCD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
JCD Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
JCD Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
JCD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
JCD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
JCD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
JCD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
JCD Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
JCD Dual& operator*=(Dual& a, Dual b) { a = a * b; return a; }
JCD Dual& operator*=(Dual& a, double b) { a.v *= b; a.d *= b; return a; }
#undef JCD

// value / tangent accessors that also accept plain doubles
__device__ __forceinline__ double jx_val(double a) { return a; }
__device__ __forceinline__ double jx_val(Dual a) { return a.v; }

// ---- elementary functions, overloaded for double and Dual --------------------------------------------
__device__ __forceinline__ double jx_rcp(double x) { return jcm_rcp(x); }
__device__ __forceinline__ Dual jx_rcp(Dual x) { const double r = jcm_rcp(x.v); return Dual(r, -x.d * r * r); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) { return a * jx_rcp(b); }
__device__ __forceinline__ Dual operator/(Dual a, double b) { const double r = 1.0 / b; return Dual(a.v * r, a.d * r); }
__device__ __forceinline__ Dual operator/(double a, Dual b) { return a * jx_rcp(b); }

__device__ __forceinline__ double jx_exp(double x) { return jcm_exp(x); }
__device__ __forceinline__ Dual jx_exp(Dual x) { const double e = jcm_exp(x.v); return Dual(e, e * x.d); }
__device__ __forceinline__ double jx_log(double x) { return jcm_log(x); }
__device__ __forceinline__ Dual jx_log(Dual x) { return Dual(jcm_log(x.v), x.d * jcm_rcp(x.v)); }
__device__ __forceinline__ double jx_exp_t(double x, const double* tab) { return jcm_exp_t(x, tab); }
__device__ __forceinline__ Dual jx_exp_t(Dual x, const double* tab) { const double e = jcm_exp_t(x.v, tab); return Dual(e, e * x.d); }
// exp of a bounded argument (|x| < 700): no underflow clamp
__device__ __forceinline__ double jx_exp_tb(double x, const double* tab) { return jcm_exp_t<false>(x, tab); }
__device__ __forceinline__ Dual jx_exp_tb(Dual x, const double* tab) { const double e = jcm_exp_t<false>(x.v, tab); return Dual(e, e * x.d); }
__device__ __forceinline__ double jx_log_t(double x, const double* tab) { return jcm_log_t(x, tab); }
__device__ __forceinline__ Dual jx_log_t(Dual x, const double* tab) { return Dual(jcm_log_t(x.v, tab), x.d * jcm_rcp(x.v)); }
__device__ __forceinline__ double jx_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ Dual jx_sqrt(Dual x) { const double s = sqrt(x.v); return Dual(s, 0.5 * x.d / s); }
__device__ __forceinline__ double jx_sin(double x) { return jcm_sin(x); }
__device__ __forceinline__ Dual jx_sin(Dual x) {  // cos(x) = sin(x + pi/2); arguments are >= 0 on this path
  return Dual(jcm_sin(x.v), jcm_sin(x.v + 1.5707963267948966) * x.d);
}
__device__ __forceinline__ double jx_rcbrt(double x) { return jcm_rcbrt(x); }
__device__ __forceinline__ Dual jx_rcbrt(Dual x) {  // d x^(-1/3) = -1/3 x^(-4/3) = -1/3 y^4
  const double y = jcm_rcbrt(x.v);
  const double y2 = y * y;
  return Dual(y, (-1.0 / 3.0) * y2 * y2 * x.d);
}
// x^y = exp(y ln x) for a positive normal base, with this file's exp / log (error ~ |y ln x| ulp, a few 1e-16 for the
// Eisenstein-Hu constants).  libm's pow costs ~300 instructions a call, and the ~20 calls of the EH-constant thread sat in
// front of the setup kernel's first barrier: 14 % of its warp time (ncu source page, BSSY at the barrier).
__device__ __forceinline__ double jx_pow(double x, double y) { return jcm_exp(y * jcm_log(x)); }
__device__ __forceinline__ Dual jx_pow(Dual x, double y) { const double p = jcm_exp(y * jcm_log(x.v)); return Dual(p, y * p / x.v * x.d); }
__device__ __forceinline__ Dual jx_pow(Dual x, Dual y) {
  const double lx = jcm_log(x.v);
  const double p = jcm_exp(y.v * lx);
  return Dual(p, p * fma(y.d, lx, y.v * x.d / x.v));
}
// selections follow the value (the sub-gradient jax takes away from ties)
__device__ __forceinline__ double jx_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ Dual jx_max(Dual a, double b) { return a.v >= b ? a : Dual(b); }
__device__ __forceinline__ Dual jx_max(Dual a, Dual b) { return a.v >= b.v ? a : b; }
__device__ __forceinline__ double jx_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ Dual jx_min(Dual a, double b) { return a.v <= b ? a : Dual(b); }
// max(x, 0) and, for x >= 0, max(x, 1) as integer selects on the words of x: 3 ALU instructions instead of
// fmax's DSETP.MAX / FSEL / SEL / NaN-fixup / register-move sequence (8-10 issue slots, one on the FP64
// pipe); the clips of the lensing-efficiency integrand (probes.py:49, background.py:242) run 3 per point.
// NaN passes through, -0 becomes +0.
__device__ __forceinline__ double jx_clip0(double x) {
  const int hi = __double2hiint(x), keep = ~(hi >> 31);
  return __hiloint2double(hi & keep, __double2loint(x) & keep);
}
__device__ __forceinline__ Dual jx_clip0(Dual a) { return a.v >= 0.0 ? a : Dual(0.0); }
__device__ __forceinline__ double jx_floor1(double x) {  // x >= 0
  const int hi = __double2hiint(x);
  const bool small = hi < 0x3ff00000;
  return __hiloint2double(small ? 0x3ff00000 : hi, small ? 0 : __double2loint(x));
}
__device__ __forceinline__ Dual jx_floor1(Dual a) { return a.v >= 1.0 ? a : Dual(1.0); }
__device__ __forceinline__ double jx_abs(double a) { return fabs(a); }
__device__ __forceinline__ Dual jx_abs(Dual a) { return a.v >= 0.0 ? a : -a; }
__device__ __forceinline__ double jx_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ Dual jx_fma(Dual a, Dual b, Dual c) { return a * b + c; }
__device__ __forceinline__ Dual jx_fma(double a, Dual b, Dual c) { return a * b + c; }
__device__ __forceinline__ Dual jx_fma(Dual a, double b, Dual c) { return a * b + c; }
__device__ __forceinline__ Dual jx_fma(Dual a, Dual b, double c) { return a * b + c; }
__device__ __forceinline__ Dual jx_fma(double a, Dual b, double c) { return a * b + c; }
__device__ __forceinline__ Dual jx_fma(Dual a, double b, double c) { return a * b + c; }
__device__ __forceinline__ Dual jx_fma(double a, double b, Dual c) { return fma(a, b, c.v) + Dual(0.0, c.d); }

__device__ __forceinline__ double jx_shfl_xor(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ Dual jx_shfl_xor(Dual v, int o) {
  return Dual(__shfl_xor_sync(0xffffffffu, v.v, o), __shfl_xor_sync(0xffffffffu, v.d, o));
}

// ---- workspace access: value plane at p[i], tangent plane at p[i + doff] --------------------------------
template <class T> struct JxMem;
template <> struct JxMem<double> {
  static __device__ __forceinline__ double ld(const double* p, ptrdiff_t) { return *p; }
  static __device__ __forceinline__ void st(double* p, ptrdiff_t, double x) { *p = x; }
};
template <> struct JxMem<Dual> {
  static __device__ __forceinline__ Dual ld(const double* p, ptrdiff_t doff) { return Dual(p[0], p[doff]); }
  static __device__ __forceinline__ void st(double* p, ptrdiff_t doff, Dual x) { p[0] = x.v; p[doff] = x.d; }
};
