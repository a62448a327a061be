// jc_power.cu -- K3: V[n,l] = geom_n * P(k = (l+1/2)/max(chi_n,1), a_n) for every (node, ell).
//
// A thread owns one ell and NPT (= 16) consecutive Limber nodes (l fastest across lanes, so a warp shares its
// node constants and V rows are written coalesced); the ell-side table entries, the per-cosmology EH
// constants and the polynomial coefficients are loaded once per thread and reused over the NPT nodes.
// Eisenstein-Hu "eisenhu_osc" transfer (transfer.py:113-153), linear power (power.py:49-52) and
// halofit takahashi2012 (power.py:246-262) are evaluated entirely in registers.  Compared with the
// reference's expression tree the arithmetic is re-associated only (<= a few ulp):
//   * every fixed-exponent power law of k = (l+1/2)/chi_c is separable: k^p = (l+1/2)^p * chi_c^-p,
//     so q^1.08, (k/k_silk)^1.4, k^(3+n_s) and k^-3 are one multiply of a per-ell (plan / ws.ellpow)
//     and a per-node (setup kernel) table entry;
//   * (1+D2L)^beta * exp(-(y/4+y^2/8)) is one exp -- 5 exp, 3 log, 1 rcbrt, 1 sin remain per point;
//   * the ~14 divisions of T(k) and of Delta^2_Q + Delta^2_H are merged into two reciprocals;
//   * exp/log are table driven (jc_math.cuh: 2^n T[j] p4(r), 256 entries per octave; {c_j, -ln c_j} + log1p degree 7), the tables
//     are staged in shared memory once per CTA and the CTA strides over its cosmology's index space;
//     sin/rcbrt/rcp and all polynomial coefficients are constant-bank DFMA operands.
// Template on the scalar type: double (hot path) or DualN<K> (JVP pass, jc_dual.cuh).
#include <cstdlib>

#include "jc_internal.cuh"
#include "jc_dual.cuh"

namespace {

struct PowerK {
  double e1, c699, c142, c386, c18, inv54, inv52, eighth, quarter, c043, two_e1, c625, c731;
};
static __constant__ PowerK PK = {2.718281828459045 /* np.exp(1.0) */, 69.9, 14.2, 386.0, 1.8,
                                 1.0 / 5.4, 1.0 / 5.2, 0.125, 0.25, 0.43, 2.0 * 2.718281828459045, 62.5, 731.0};

// Per-cosmology constants of the transfer function (transfer.py:47-136), loaded once per thread.
template <class T> struct EhK {
  T inv13keq, c14ac, b18, shd, alpha_b, fb, fc, bnode3, bb3, alpha_g, omh_t27;
};
template <class T> __device__ __forceinline__ EhK<T> eh_load(const double* scp, ptrdiff_t doff) {
  auto SC = [&](int f) { return JxMem<T>::ld(scp + f, doff); };
  EhK<T> E;
  E.inv13keq = SC(JC_SCAL_INV13KEQ); E.c14ac = SC(JC_SCAL_C14_ALPHA_C); E.b18 = PK.c18 * SC(JC_SCAL_BETA_C);
  E.shd = SC(JC_SCAL_SH_D); E.alpha_b = SC(JC_SCAL_ALPHA_B); E.fb = SC(JC_SCAL_FB); E.fc = SC(JC_SCAL_FC);
  const T bnode = SC(JC_SCAL_BETA_NODE), bb = SC(JC_SCAL_BETA_B);
  E.bnode3 = bnode * bnode * bnode; E.bb3 = bb * bb * bb;
  E.alpha_g = SC(JC_SCAL_ALPHA_GAMMA); E.omh_t27 = SC(JC_SCAL_OMH_T27);  // no-wiggle fit only
  return E;
}

// T(k): Eisenstein & Hu with wiggles (transfer.py:113-153) or the no-wiggle fit (transfer.py:92-105).
// q108 = q^1.08 and ks14 = (k / k_silk)^1.4 come from the caller (separable tables in the point kernel, exp in the
// table build); the ~14 divisions are merged into one reciprocal.
template <class T, bool NOWIG>
__device__ __forceinline__ T eh_point(const EhK<T>& E, T k, T q108, T ks14, const double* __restrict__ s_tab) {
  if constexpr (NOWIG) {
    const T ks43 = PK.c043 * (k * E.shd);
    const T k2 = ks43 * ks43;
    const T q = k * jx_rcp(E.omh_t27 * (E.alpha_g + (JCK.one - E.alpha_g) * jx_rcp(k2 * k2 + JCK.one)));
    const T L = jx_log_t(PK.c18 * q + PK.two_e1, s_tab);
    const T Wn = PK.c625 * q + JCK.one;                  // C = 14.2 + 731/Wn
    const T LW = L * Wn;
    return LW * jx_rcp(LW + (PK.c142 * Wn + PK.c731) * (q * q));
  } else {
    const T q = k * E.inv13keq;
    const T q2 = q * q;
    const T W = PK.c699 * q108 + JCK.one;  // 1 + 69.9 q^1.08
    const T U1 = PK.c142 * W + PK.c386;    // C(alpha=1) W
    const T U2 = E.c14ac * W + PK.c386;    // C(alpha_c) W
    const T L1 = jx_log_t(E.b18 * q + PK.e1, s_tab);
    const T L2 = jx_log_t(PK.c18 * q + PK.e1, s_tab);
    const T L1W = L1 * W, L2W = L2 * W;
    const T N1 = U1 * q2 + L1W;  // T~(k,1,beta_c)       = L1W / N1
    const T N2 = U2 * q2 + L1W;  // T~(k,alpha_c,beta_c) = L1W / N2
    const T N3 = U1 * q2 + L2W;  // T~(k,1,1)            = L2W / N3
    const T ks = k * E.shd;
    const T x54 = ks * PK.inv54;
    const T x54_2 = x54 * x54;
    const T Fm1 = x54_2 * x54_2;  // f = 1/(1+Fm1)
    // Tc = f T1 + (1-f) T2 = L1W (N2 + Fm1 N1) / ((1+Fm1) N1 N2)
    const T numC = L1W * (Fm1 * N1 + N2);
    const T denC = (JCK.one + Fm1) * (N1 * N2);
    const T ks2 = ks * ks, ks3 = ks2 * ks;
    const T arg = ks2 * jx_rcbrt(ks3 + E.bnode3);  // k s~ = ks^2 / cbrt(ks^3 + beta_node^3)
    const T x52 = ks * PK.inv52;
    const T X52 = x52 * x52 + JCK.one;
    const T BB = ks3 + E.bb3;  // 1/(1+(beta_b/ks)^3) = ks^3 / BB
    const T silk = jx_exp_t(-ks14, s_tab);  // exp(-(k/k_silk)^1.4)
    // Tb = [T3/X52 + alpha_b ks^3/BB silk] sin(arg)/arg
    const T N3X = N3 * X52;
    const T numB = (L2W * BB + E.alpha_b * ks3 * silk * N3X) * jx_sin(arg);
    const T denB = N3X * BB * arg;
    return ((E.fb * numB) * denC + E.fc * numC * denB) * jx_rcp(denB * denC);
  }
}

// NPT: Limber nodes per thread; MINB: CTAs per SM the register allocation must allow; NOWIG: the
// no-wiggle Eisenstein-Hu fit (transfer.py:99-105) instead of the default "eisenhu_osc".
template <class T, int NPT, int MINB, bool NOWIG>
__global__ void __launch_bounds__(256, MINB) jc_power_kernel(JcDevPlan pl, Ws ws, unsigned inv_L) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
  __shared__ __align__(16) double s_tab[JCM_TAB_DOUBLES];
  for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += 256) s_tab[i] = pl.math_tab[i];
  __syncthreads();
  const ptrdiff_t doff = ws.doff;
  const int c = blockIdx.y;
  const double* scp = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  const EhK<T> E = eh_load<T>(scp, doff);  // cosmology side (transfer.py:47-136)
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD;
#define NODE(f) JxMem<T>::ld(nd + (f)*JC_NA_PAD + n, doff)

  for (unsigned idx = blockIdx.x * 256 + threadIdx.x; idx < (unsigned)(NGRP * pl.L); idx += gridDim.x * 256) {
    // idx / L: multiply-high by inv_L = ceil(2^32 / L) is exact while 513 L^2 < 2^32 (inv_L = 0 otherwise)
    const int grp = inv_L ? (int)__umulhi(idx, inv_L) : (int)(idx / (unsigned)pl.L);
    const int l = (int)idx - grp * pl.L;
    // ell side
    const double lnl = pl.lnellp5[l], lp5 = pl.ellp5[l], l108 = pl.ell108[l], l14 = pl.ell14[l], lm3 = pl.ellm3[l];
    const T lpns = JxMem<T>::ld(ws.ellpow + (size_t)c * pl.Lpad + l, doff);
    const int n0 = grp * NPT;
    const int n1 = min(n0 + NPT, JC_NA);
    double* vout = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l;

#pragma unroll 1
    for (int n = n0; n < n1; ++n) {
      const T lnk = lnl - NODE(JC_NODE_LNCHIC);
      const T k = lp5 * NODE(JC_NODE_INVCHIC);  // angular_cl.py:73
      const T Tk = eh_point<T, NOWIG>(E, k, NOWIG ? T(0.0) : T(l108 * NODE(JC_NODE_NQ108)), NOWIG ? T(0.0) : T(l14 * NODE(JC_NODE_NSILK)), s_tab);
      // ---- Delta^2_L = k^3 P_lin / (2 pi^2)  (power.py:49-52, :250) -----------------------------------
      const T d2l = lpns * NODE(JC_NODE_NAMP) * (Tk * Tk);
      T d2;
      if (pl.nonlinear) {  // halofit, takahashi2012 (power.py:246-262)
        const T y = k * NODE(JC_NODE_RNL);
        const T lny = lnk - NODE(JC_NODE_LNKNL);
        const T y2 = y * y;
        // Delta^2_Q = D2L (1+D2L)^beta / (1 + alpha D2L) exp(-(y/4 + y^2/8))
        const T Nq = d2l * jx_exp_t(NODE(JC_NODE_BETA) * jx_log_t(JCK.one + d2l, s_tab) - (y2 * PK.eighth + PK.quarter * y), s_tab);
        const T Dq = NODE(JC_NODE_ALPHA) * d2l + JCK.one;
        // y^(3 f1), y^(f2), c f3 y^(3-gamma): |ln y| < 30 and exponents of order one -- bounded arguments, no clamp
        const T ye1 = jx_exp_tb(NODE(JC_NODE_E1) * lny, s_tab);
        const T ye2 = jx_exp_tb(NODE(JC_NODE_E2) * lny, s_tab);
        const T cfy = jx_exp_tb(NODE(JC_NODE_P3) * (NODE(JC_NODE_LNCF) + lny), s_tab);
        const T Nh = NODE(JC_NODE_AN) * ye1 * y2;
        T ynu = y2 + NODE(JC_NODE_NU);  // 1/(1 + mu/y + nu/y^2) = y^2 / (y^2 + mu y + nu)
        if (pl.nonlinear == JC_PK_HALOFIT_SMITH2003) ynu = ynu + NODE(JC_NODE_MU) * y;  // mu = 0 in takahashi2012
        const T Dh = (NODE(JC_NODE_BN) * ye2 + JCK.one + cfy) * ynu;
        d2 = (Nq * Dh + Nh * Dq) * jx_rcp(Dq * Dh);  // Delta^2_Q + Delta^2_H
      } else {
        d2 = d2l;
      }
      // P = 2 pi^2 / k^3 * Delta^2 ;  V = P * geom = Delta^2 (l+1/2)^-3 * [geom 2 pi^2 chi_c^3]
      JxMem<T>::st(vout + (size_t)n * pl.Lpad, doff, d2 * lm3 * NODE(JC_NODE_GK));
    }
  }
#undef NODE
}


// ---------------------------------------------------------------------------------------------------------------------
// K3 with the transfer function tabulated per cosmology (the hot-path default for double, non-grid plans with >= 32 ell).
//
// T(k) depends on k only, and the 51,300 (ell, node) points of a cosmology sample it densely: ~80 of the exact kernel's
// 186 FP64 instructions per point re-evaluate the same smooth function.  Here one CTA (512 threads, two per SM) owns a
// cosmology: it evaluates T(k) with the exact formula at NT nodes uniform in ln k over the cosmology's own range
// [ln(l_min+1/2) - ln chi_max, ln(l_max+1/2)] (spacing h = range / (NT - 7) = 0.0023 at the bench configuration),
// derives the node slopes h T'(k_j) by 4th-order central differences, keeps {T, h T'} pairs in shared memory (96 KB) and
// evaluates every point by cubic Hermite interpolation (2 LDS.128 + 11 FP64 instead of 2 log, 1 exp, 1 sin, 1 rcbrt and
// ~45 multiply-adds).  The halofit part uses the reduced-degree exp / log of jc_math.cuh.
// Accuracy: the interpolation error is largest where the baryon wiggles are densest in ln k (k ~ 0.4 h/Mpc: 8e-10 on a single V entry);
// on C_ell, an integral over 513 nodes, it is <= 6e-11 (oracle experiment over the config-5 box, profiles/r02_power_tab.md) against the
// path's 1e-6 bar.  The exact kernel above stays the one behind grid plans, JVP passes, L < 32 and JC_POWER_EXACT=1,
// and a cosmology whose range would need h > 0.004 takes the exact formula inside this kernel (warp-uniform branch).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TAB_NT = 6144;           // table nodes (3 pad nodes on either side of the range)
constexpr int TAB_THREADS = 512;
constexpr double TAB_H_MAX = 0.004;

// The (ell, node) points of one cosmology for the CTA of jc_power_tab_kernel.  TAB: T(k) by cubic Hermite interpolation
// in the shared-memory table; !TAB: the exact formula (cosmologies whose ln k range is too wide for the table) -- a
// separate instantiation so that the EH constants are not live in the tabulated loop.
template <int NPT, bool NOWIG, bool TAB, int NTHREADS = 512>
__device__ __forceinline__ void tab_points(const JcDevPlan& pl, const Ws& ws, int c, unsigned inv_L, const EhK<double>& E,
                                           const double* __restrict__ nd, const double2* __restrict__ tk,
                                           const double* __restrict__ s_tab, double x0, double inv_h) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
#define NODE(f) nd[(f)*JC_NA_PAD + n]
  for (unsigned idx = threadIdx.x; idx < (unsigned)(NGRP * pl.L); idx += NTHREADS) {
    const int grp = inv_L ? (int)__umulhi(idx, inv_L) : (int)(idx / (unsigned)pl.L);
    const int l = (int)idx - grp * pl.L;
    const double lnl = pl.lnellp5[l], lp5 = pl.ellp5[l], lm3 = pl.ellm3[l];
    const double lpns = ws.ellpow[(size_t)c * pl.Lpad + l];
    const int n0 = grp * NPT;
    const int n1 = min(n0 + NPT, JC_NA);
    double* vout = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l;
#pragma unroll 1
    for (int n = n0; n < n1; ++n) {
      const double lnk = lnl - NODE(JC_NODE_LNCHIC);
      const double k = lp5 * NODE(JC_NODE_INVCHIC);  // angular_cl.py:73
      double Tk;
      if constexpr (TAB) {
        const double u = (lnk - x0) * inv_h;
        int i = __double2int_rz(u);
        i = max(2, min(i, TAB_NT - 4));
        const double t = u - (double)i;
        const double2 a = tk[i], b = tk[i + 1];
        const double D = b.x - a.x;
        const double c3 = (a.y + b.y) - (D + D);  // cubic Hermite in t: f0 + t (d0 + t (c2 + t c3))
        const double c2 = (D - a.y) - c3;
        Tk = fma(t, fma(t, fma(t, c3, c2), a.y), a.x);
      } else {
        Tk = eh_point<double, NOWIG>(E, k, NOWIG ? 0.0 : pl.ell108[l] * NODE(JC_NODE_NQ108), NOWIG ? 0.0 : pl.ell14[l] * NODE(JC_NODE_NSILK), s_tab);
      }
      const double d2l = lpns * NODE(JC_NODE_NAMP) * (Tk * Tk);
      double d2;
      if (pl.nonlinear) {  // halofit (power.py:246-262), same association as the exact kernel
        const double y = k * NODE(JC_NODE_RNL);
        const double lny = lnk - NODE(JC_NODE_LNKNL);
        const double y2 = y * y;
        const double Nq = d2l * jcm_exp_t3(NODE(JC_NODE_BETA) * jcm_log_t4(JCK.one + d2l, s_tab) - (y2 * PK.eighth + PK.quarter * y), s_tab);
        const double Dq = NODE(JC_NODE_ALPHA) * d2l + JCK.one;
        const double ye1 = jcm_exp_t3<false>(NODE(JC_NODE_E1) * lny, s_tab);
        const double ye2 = jcm_exp_t3<false>(NODE(JC_NODE_E2) * lny, s_tab);
        const double cfy = jcm_exp_t3<false>(NODE(JC_NODE_P3) * (NODE(JC_NODE_LNCF) + lny), s_tab);
        const double Nh = NODE(JC_NODE_AN) * ye1 * y2;
        double ynu = y2 + NODE(JC_NODE_NU);
        if (pl.nonlinear == JC_PK_HALOFIT_SMITH2003) ynu = ynu + NODE(JC_NODE_MU) * y;
        const double Dh = (NODE(JC_NODE_BN) * ye2 + JCK.one + cfy) * ynu;
        d2 = (Nq * Dh + Nh * Dq) * jcm_rcp(Dq * Dh);
      } else {
        d2 = d2l;
      }
      vout[(size_t)n * pl.Lpad] = d2 * lm3 * NODE(JC_NODE_GK);
    }
  }
#undef NODE
}

template <int NPT, bool NOWIG>
__global__ void __launch_bounds__(TAB_THREADS, 2) jc_power_tab_kernel(JcDevPlan pl, Ws ws, unsigned inv_L, double lnl_min,
                                                                       double lnl_max) {
  extern __shared__ __align__(16) double smem_p[];
  double2* tk = reinterpret_cast<double2*>(smem_p);  // [TAB_NT] {T(k_j), h dT/dlnk(k_j)}
  double* s_tab = smem_p + 2 * TAB_NT;
  for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += TAB_THREADS) s_tab[i] = pl.math_tab[i];
  const int c = blockIdx.x;
  const double* scp = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  const EhK<double> E = eh_load<double>(scp, 0);
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD;
  // ln k range of this cosmology: chi decreases with the node index, node 0 = a_min holds the largest chi
  const double lo = lnl_min - nd[JC_NODE_LNCHIC * JC_NA_PAD + 0], hi = lnl_max - nd[JC_NODE_LNCHIC * JC_NA_PAD + JC_NA - 1];
  const double h = (hi - lo) * (1.0 / (TAB_NT - 7));
  const bool use_tab = h <= TAB_H_MAX;
  const double x0 = lo - 3.0 * h, inv_h = 1.0 / h;
  __syncthreads();
  if (use_tab) {
    const double ln13keq = scp[JC_SCAL_LN13KEQ], lnksilk = scp[JC_SCAL_LNKSILK];
    for (int j = threadIdx.x; j < TAB_NT; j += TAB_THREADS) {
      const double lnk = fma((double)j, h, x0);
      const double k = jcm_exp_t(lnk, s_tab);
      double q108 = 0.0, ks14 = 0.0;
      if (!NOWIG) {
        q108 = jcm_exp_t(1.08 * (lnk - ln13keq), s_tab);
        ks14 = jcm_exp_t(1.4 * (lnk - lnksilk), s_tab);
      }
      tk[j].x = eh_point<double, NOWIG>(E, k, q108, ks14, s_tab);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < TAB_NT; j += TAB_THREADS) {
      double d = 0.0;
      if (j >= 2 && j < TAB_NT - 2)
        d = (8.0 * (tk[j + 1].x - tk[j - 1].x) - (tk[j + 2].x - tk[j - 2].x)) * (1.0 / 12.0);
      tk[j].y = d;
    }
    __syncthreads();
  }
  if (use_tab) tab_points<NPT, NOWIG, true>(pl, ws, c, inv_L, E, nd, tk, s_tab, x0, inv_h);
  else tab_points<NPT, NOWIG, false>(pl, ws, c, inv_L, E, nd, tk, s_tab, x0, inv_h);
}


// ---------------------------------------------------------------------------------------------------------------------
// K3, tabulated, second form: one CTA of 1024 threads per cosmology with EVERYTHING the point loop touches in shared
// memory.  ncu on jc_power_tab_kernel (profiles/r02_ncu_summary.md): issue slots 48 % busy, 1.0 eligible warp per
// scheduler; stalls = long scoreboard 6.1 warps per issue (the ~17 per-node table entries a point reads come from L2: with
// 2 x 102 KB of shared memory carved out, 28 KB of L1 are left), wait 3.1, short scoreboard 2.8; shared-memory wavefronts
// 59 per warp-point, 60 % of them bank conflicts of the random table gathers.  Here
//   * a thread owns ONE ell for the whole kernel (row r = tid / L of floor(1024 / L) rows): its five ell-side values sit
//     in registers;
//   * the Limber nodes are walked in blocks of nb <= 64; the 15 per-node values of a block, already combined into the
//     forms the point needs (table coordinate offset, y / (l+1/2), ln y offset, ...), are staged in shared memory one
//     block ahead (loads at the top of a block, stores after its points, one __syncthreads per block): warp-uniform LDS
//     instead of L2 round trips;
//   * the exp / log tables are replicated per lane slot ("lane-private banks": entry j of slot c at word j * 16 + c), so
//     a gather with 32 different indices is conflict free: 2 wavefronts per LDS.64 instead of ~6, 4 per LDS.128 instead
//     of ~10; exp uses 128 entries per octave with the degree-3 polynomial (|r| <= ln2/256, truncation 2e-12);
//   * when ln(ell + 1/2) is uniformly spaced (plan: lnl_step > 0, e.g. np.logspace) the table spacing h is snapped so that
//     neighbouring lanes are an ODD number of entries apart: the two LDS.128 of the Hermite gather then touch every
//     16-byte bank group once per quarter warp (4 wavefronts each instead of ~10).
// Used when floor(1024 / L) * L >= 0.9 * 1024 (else jc_power_tab_kernel).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int T2_THREADS = 1024;
constexpr int T2_NB = 64;   // nodes per block (<=)
constexpr int T2_NF = 15;   // staged values per node
constexpr int T2_EXPN = 128;
enum { F_UOFF = 0, F_YF, F_LNY0, F_NAMP, F_BETA, F_ALPHA, F_E1, F_E2, F_P3, F_P3LNCF, F_AN, F_NU, F_BN, F_GK, F_MU };

struct JcMathL {
  double k128, magic, l128;  // 128/ln2, 1.5*2^52, ln2/128
};
static __constant__ JcMathL JCL = {1.4426950408889634074 * T2_EXPN, 6755399441055744.0, 6.93147180559945309417e-01 / T2_EXPN};

// exp(x), |x| < 700 unless CLAMP; lane-private table: entry j of slot c at e[j * 16 + c].  `e_slot_addr` is the 32-bit
// shared-memory address of the lane's slot: table address and exponent insertion are two integer instructions each
// (LOP3 + IMAD, SHF + LEA) instead of the seven the indexed C++ form compiled to.
template <bool CLAMP>
__device__ __forceinline__ double exp_lane(double x, unsigned e_slot_addr) {
  if (CLAMP) x = jcm_clamp_exp_arg(x);
  const double kd = fma(x, JCL.k128, JCL.magic);
  const int k = __double2loint(kd);
  const double kf = kd - JCL.magic;
  const double r = fma(kf, -JCL.l128, x);
  double p = fma(JCT.e[1], r, JCT.e[0]);
  p = fma(p, r, JCK.one);
  p = fma(p, r, JCK.one);
  double t;
  asm("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(e_slot_addr + (unsigned)(k & (T2_EXPN - 1)) * 128u));
  p *= t;
  return __hiloint2double(__double2hiint(p) + ((k >> 7) << 20), __double2loint(p));
}
// log(x), x >= 1 normal; lane-private table of {c_j, -ln c_j}: entry j of slot c at l[(j * 8 + c)]
__device__ __forceinline__ double log_lane(double x, unsigned l_slot_addr) {
  const int hx = __double2hiint(x);
  const int e = (hx >> 20) - 1023;
  const double m = __hiloint2double((hx & 0x000fffff) | 0x3ff00000, __double2loint(x));
  double cx, cy;
  // j = top 7 mantissa bits = (hx >> 13) & 127, entry stride 128 bytes: (hx >> 6) & 0x3f80
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cx), "=d"(cy) : "r"(l_slot_addr + (((unsigned)hx >> 6) & 0x3f80u)));
  const double r = fma(m, cx, -JCK.one);
  double p = fma(JCT.l[2], r, JCT.l[1]);
  p = fma(p, r, JCT.l[0]);
  p = fma(p * r, r, r);
  return fma((double)e, JCK.ln2_hi, cy) + p;
}

template <bool NOWIG>
__global__ void __launch_bounds__(T2_THREADS, 1) jc_power_tab2_kernel(JcDevPlan pl, Ws ws, int nb, int rows) {
  extern __shared__ __align__(16) double smem_p[];
  double2* tk = reinterpret_cast<double2*>(smem_p);                 // [TAB_NT] {T(k_j), h dT/dlnk(k_j)}
  double* s_tab = smem_p + 2 * TAB_NT;                               // build phase: jc_math tables
  double* e_l = s_tab + JCM_TAB_DOUBLES;                             // [128][16]
  double2* l_l = reinterpret_cast<double2*>(e_l + T2_EXPN * 16);     // [128][8]
  double* nf = e_l + T2_EXPN * 16 + 128 * 8 * 2;                     // [2][T2_NF][T2_NB]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < JCM_TAB_DOUBLES; i += T2_THREADS) s_tab[i] = pl.math_tab[i];
  for (int i = tid; i < T2_EXPN * 16; i += T2_THREADS) e_l[i] = pl.math_tab[JCM_TAB_EXP + (i >> 4) * (JCM_EXP_N / T2_EXPN)];
  for (int i = tid; i < 128 * 8; i += T2_THREADS)
    l_l[i] = make_double2(pl.math_tab[JCM_TAB_LOG + 2 * (i >> 3)], pl.math_tab[JCM_TAB_LOG + 2 * (i >> 3) + 1]);
  const int c = blockIdx.x;
  const double* scp = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD;
  const double lo = pl.lnl_min - nd[JC_NODE_LNCHIC * JC_NA_PAD + 0], hi = pl.lnl_max - nd[JC_NODE_LNCHIC * JC_NA_PAD + JC_NA - 1];
  double h = (hi - lo) * (1.0 / (TAB_NT - 7));
  if (pl.lnl_step > 0.0) {  // snap: neighbouring ell an odd number of table entries apart
    int m = (int)(pl.lnl_step / h);
    m -= 1 - (m & 1);
    if (m >= 1) h = pl.lnl_step / (double)m;
  }
  const bool use_tab = h <= TAB_H_MAX;
  const double x0 = lo - 3.0 * h, inv_h = 1.0 / h;
  __syncthreads();
  if (use_tab) {
    const EhK<double> E = eh_load<double>(scp, 0);
    const double ln13keq = scp[JC_SCAL_LN13KEQ], lnksilk = scp[JC_SCAL_LNKSILK];
    for (int j = tid; j < TAB_NT; j += T2_THREADS) {
      const double lnk = fma((double)j, h, x0);
      const double k = jcm_exp_t(lnk, s_tab);
      double q108 = 0.0, ks14 = 0.0;
      if (!NOWIG) {
        q108 = jcm_exp_t(1.08 * (lnk - ln13keq), s_tab);
        ks14 = jcm_exp_t(1.4 * (lnk - lnksilk), s_tab);
      }
      tk[j].x = eh_point<double, NOWIG>(E, k, q108, ks14, s_tab);
    }
    __syncthreads();
    for (int j = tid; j < TAB_NT; j += T2_THREADS) {
      double d = 0.0;
      if (j >= 2 && j < TAB_NT - 2)
        d = (8.0 * (tk[j + 1].x - tk[j - 1].x) - (tk[j + 2].x - tk[j - 2].x)) * (1.0 / 12.0);
      tk[j].y = d;
    }
  } else {  // ln k range too wide for the table: exact formula at every point (global-memory node tables)
    const EhK<double> E = eh_load<double>(scp, 0);
    const unsigned inv_L = (pl.L >= 2 && pl.L <= 2048) ? (unsigned)((0x100000000ull + pl.L - 1) / pl.L) : 0u;
    tab_points<8, NOWIG, false, T2_THREADS>(pl, ws, c, inv_L, E, nd, tk, s_tab, x0, inv_h);
    return;
  }
  // ---- point loop -------------------------------------------------------------------------------------------------
  const int L = pl.L;
  const int row = tid / L, l = tid - row * L;
  const bool active = row < rows;
  double lnl = 0.0, lnl_ih = 0.0, lp5 = 0.0, lpns = 0.0, lm3 = 0.0;
  if (active) {
    lnl = pl.lnellp5[l]; lp5 = pl.ellp5[l]; lm3 = pl.ellm3[l];
    lpns = ws.ellpow[(size_t)c * pl.Lpad + l];
    lnl_ih = lnl * inv_h;
  }
  const unsigned e_slot = (unsigned)__cvta_generic_to_shared(e_l + (lane & 15));
  const unsigned l_slot = (unsigned)__cvta_generic_to_shared(l_l + (lane & 7));
  const bool smith = pl.nonlinear == JC_PK_HALOFIT_SMITH2003;
  const bool halofit = pl.nonlinear != 0;
  double* vbase = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l;
  // staging of a node block: thread t < nb * T2_NF loads and combines one value (field f = t / nb, node j = t % nb)
  auto stage_load = [&](int blk) -> double {
    const int f = tid / nb, j = tid - f * nb, n = blk * nb + j;
    if (f >= T2_NF || n >= JC_NA) return 0.0;
    auto N = [&](int fld) { return nd[fld * JC_NA_PAD + n]; };
    switch (f) {
      case F_UOFF: return -(N(JC_NODE_LNCHIC) + x0) * inv_h;
      case F_YF: return N(JC_NODE_INVCHIC) * N(JC_NODE_RNL);
      case F_LNY0: return N(JC_NODE_LNCHIC) + N(JC_NODE_LNKNL);
      case F_NAMP: return N(JC_NODE_NAMP);
      case F_BETA: return N(JC_NODE_BETA);
      case F_ALPHA: return N(JC_NODE_ALPHA);
      case F_E1: return N(JC_NODE_E1);
      case F_E2: return N(JC_NODE_E2);
      case F_P3: return N(JC_NODE_P3);
      case F_P3LNCF: return N(JC_NODE_P3) * N(JC_NODE_LNCF);
      case F_AN: return N(JC_NODE_AN);
      case F_NU: return N(JC_NODE_NU);
      case F_BN: return N(JC_NODE_BN);
      case F_GK: return N(JC_NODE_GK);
      default: return N(JC_NODE_MU);
    }
  };
  const int nblk = (JC_NA + nb - 1) / nb;
  if (tid < nb * T2_NF) nf[(tid / nb) * T2_NB + (tid % nb)] = stage_load(0);
  __syncthreads();  // table slopes + block 0
  for (int blk = 0; blk < nblk; ++blk) {
    const double* F = nf + (blk & 1) * (T2_NF * T2_NB);
    double staged = 0.0;
    if (blk + 1 < nblk) staged = stage_load(blk + 1);
    const int n_lo = blk * nb, n_cnt = min(nb, JC_NA - n_lo);
    if (active) {
      double* vp = vbase + (size_t)(n_lo + row) * pl.Lpad;
      const size_t vstep = (size_t)rows * pl.Lpad;
#pragma unroll 1
      for (int j = row; j < n_cnt; j += rows, vp += vstep) {
        const double u = lnl_ih + F[F_UOFF * T2_NB + j];
        int i = __double2int_rz(u);
        i = max(2, min(i, TAB_NT - 4));
        const double t = u - (double)i;
        const double2 a = tk[i], b = tk[i + 1];
        const double D = b.x - a.x;
        const double c3 = (a.y + b.y) - (D + D);
        const double c2 = (D - a.y) - c3;
        const double Tk = fma(t, fma(t, fma(t, c3, c2), a.y), a.x);
        const double d2l = lpns * F[F_NAMP * T2_NB + j] * (Tk * Tk);
        double d2 = d2l;
        if (halofit) {  // power.py:246-262, same association as the exact kernel
          const double y = lp5 * F[F_YF * T2_NB + j];
          const double lny = lnl - F[F_LNY0 * T2_NB + j];
          const double y2 = y * y;
          const double Nq = d2l * exp_lane<true>(F[F_BETA * T2_NB + j] * log_lane(JCK.one + d2l, l_slot) - (y2 * PK.eighth + PK.quarter * y), e_slot);
          const double Dq = F[F_ALPHA * T2_NB + j] * d2l + JCK.one;
          const double ye1 = exp_lane<false>(F[F_E1 * T2_NB + j] * lny, e_slot);
          const double ye2 = exp_lane<false>(F[F_E2 * T2_NB + j] * lny, e_slot);
          const double cfy = exp_lane<false>(fma(F[F_P3 * T2_NB + j], lny, F[F_P3LNCF * T2_NB + j]), e_slot);
          const double Nh = F[F_AN * T2_NB + j] * ye1 * y2;
          double ynu = y2 + F[F_NU * T2_NB + j];
          if (smith) ynu = fma(F[F_MU * T2_NB + j], y, ynu);
          const double Dh = (F[F_BN * T2_NB + j] * ye2 + JCK.one + cfy) * ynu;
          d2 = (Nq * Dh + Nh * Dq) * jcm_rcp(Dq * Dh);
        }
        *vp = d2 * lm3 * F[F_GK * T2_NB + j];
      }
    }
    if (blk + 1 < nblk && tid < nb * T2_NF) nf[((blk + 1) & 1) * (T2_NF * T2_NB) + (tid / nb) * T2_NB + (tid % nb)] = staged;
    __syncthreads();
  }
}

void launch_power_tab2(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  const size_t smem = (size_t)(2 * TAB_NT + JCM_TAB_DOUBLES + T2_EXPN * 16 + 128 * 8 * 2 + 2 * T2_NF * T2_NB) * sizeof(double);
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, {
    cudaFuncSetAttribute(jc_power_tab2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(jc_power_tab2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  const int rows = T2_THREADS / pl.L;
  int nb = (T2_NB / rows) * rows;  // whole rounds of `rows` nodes per block
  if (nb < rows) nb = rows;
  if (pl.transfer == JC_TF_EISENSTEIN_HU_NOWIGGLE) jc_power_tab2_kernel<true><<<chunk, T2_THREADS, smem, s>>>(pl, ws, nb, rows);
  else jc_power_tab2_kernel<false><<<chunk, T2_THREADS, smem, s>>>(pl, ws, nb, rows);
}

template <int NPT>
void launch_power_tab(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  const unsigned inv_L = (pl.L >= 2 && pl.L <= 2048) ? (unsigned)((0x100000000ull + pl.L - 1) / pl.L) : 0u;
  const size_t smem = (size_t)(2 * TAB_NT + JCM_TAB_DOUBLES) * sizeof(double);
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, {
    cudaFuncSetAttribute(jc_power_tab_kernel<NPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(jc_power_tab_kernel<NPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (pl.transfer == JC_TF_EISENSTEIN_HU_NOWIGGLE)
    jc_power_tab_kernel<NPT, true><<<chunk, TAB_THREADS, smem, s>>>(pl, ws, inv_L, pl.lnl_min, pl.lnl_max);
  else
    jc_power_tab_kernel<NPT, false><<<chunk, TAB_THREADS, smem, s>>>(pl, ws, inv_L, pl.lnl_min, pl.lnl_max);
}

// split: CTAs per cosmology (each strides over the index space)
template <class T, int NPT, int MINB>
void launch_power_cfg(const JcDevPlan& pl, const Ws& ws, int chunk, int split, cudaStream_t s) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
  const unsigned inv_L = (pl.L >= 2 && pl.L <= 2048) ? (unsigned)((0x100000000ull + pl.L - 1) / pl.L) : 0u;
  const int full = (NGRP * pl.L + 255) / 256;
  const int gx = split < full ? split : full;
  if (pl.transfer == JC_TF_EISENSTEIN_HU_NOWIGGLE)
    jc_power_kernel<T, NPT, MINB, true><<<dim3(gx, chunk), 256, 0, s>>>(pl, ws, inv_L);
  else
    jc_power_kernel<T, NPT, MINB, false><<<dim3(gx, chunk), 256, 0, s>>>(pl, ws, inv_L);
}

}  // namespace

void jc_launch_power(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  // jc_set_option("power_exact", 1) / JC_POWER_EXACT=1: the exact-formula kernel everywhere (A/B runs, stage tests)
  if (!g_jc_power_exact && !pl.grid_mode && pl.L >= 32 && ws.doff == 0) {  // tabulated transfer function (see jc_power_tab_kernel)
    static const int first_form = [] { const char* e = getenv("JC_POWER_TAB_NPT"); return e && atoi(e) < 0; }();  // A/B knob
    const int rows = T2_THREADS / pl.L;
    if (!first_form && rows >= 1 && rows * pl.L * 10 >= T2_THREADS * 9 && pl.L <= T2_THREADS) launch_power_tab2(pl, ws, chunk, s);
    else launch_power_tab<8>(pl, ws, chunk, s);
    return;
  }
  // exact kernel; fastest shape of the round-1 sweep (profiles/r01_tuning.md: 17 shapes, 2..64 nodes per thread, 48..128
  // registers): 16 nodes per thread (the ell-side loads and index arithmetic amortise over 16 points), 64 registers,
  // 8 CTAs per cosmology: 6.13 ms per 8192 cosmologies
  launch_power_cfg<double, 16, 4>(pl, ws, chunk, 8, s);
}

// JVP passes: exact kernel on DualN<ntan>.  One direction: 16 nodes per thread at 3 CTAs / SM (7-tangent batch JVP
// 38.6 -> 31.0 ms per 1024 cosmologies against one node per thread, profiles/r01_tuning.md).  Tangent groups hold
// (1 + ntan) x the live values: fewer nodes per thread, registers up to the 255 cap.
void jc_launch_power_jvp(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s) {
  switch (ntan) {
    case 2: launch_power_cfg<DualN<2>, 8, 2>(pl, ws, chunk, 8, s); break;
    case 3: launch_power_cfg<DualN<3>, 8, 1>(pl, ws, chunk, 8, s); break;
    case 4: launch_power_cfg<DualN<4>, 8, 1>(pl, ws, chunk, 8, s); break;
    default: launch_power_cfg<Dual, 16, 3>(pl, ws, chunk, 8, s); break;
  }
}
