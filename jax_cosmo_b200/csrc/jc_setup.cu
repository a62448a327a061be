// jc_setup.cu -- K1: per-cosmology setup kernel (background tables, EH constants, sigma8 norm, halofit).
// Template on the scalar type: double (hot path) or DualN<K> (value + K tangents, JVP passes).
#include "jc_internal.cuh"
#include "jc_dual.cuh"

namespace {

constexpr double HF_HALF_LN_CUT = 2.1910133173369406;  // ln sqrt(80): halofit Gaussian-window truncation

template <class T> struct M2 { T a, b, c, d; };  // [[a b][c d]]
template <class T> __device__ __forceinline__ M2<T> mul(const M2<T>& x, const M2<T>& y) {
  return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
template <class T> __device__ __forceinline__ M2<T> eye_plus(double s, const M2<T>& x) {  // I + s*x
  return {1.0 + s * x.a, s * x.b, s * x.c, 1.0 + s * x.d};
}

template <class T> struct Bg { T Om, Ok, Ode, w0, wa; };

// E^2(a), background.py:122-126 (also returns the dark-energy term)
template <class T> __device__ __forceinline__ T esqr(const Bg<T>& c, double a, double lna, T* de_term) {
  const double ia = 1.0 / a, ia2 = ia * ia;
  const T fde = -3.0 * (1.0 + c.w0 + c.wa) * lna + 3.0 * c.wa * (a - 1.0);  // background.py:90
  const T de = c.Ode * jx_exp(fde);
  *de_term = de;
  return c.Om * (ia2 * ia) + c.Ok * ia2 + de;
}

// T(k) of transfer.py:113-153 ("eisenhu_osc") at a fixed-grid k (sigma8 and halofit nodes)
template <class T> __device__ __forceinline__ T eh_transfer(const T* __restrict__ s, double k, double lnk, int kind) {
  const double E1 = 2.718281828459045;  // np.exp(1.0)
  if (kind == JC_TF_EISENSTEIN_HU_NOWIGGLE) {  // transfer.py:92-105
    const T ks43 = 0.43 * k * s[JC_SCAL_SH_D];
    const T k2 = ks43 * ks43;
    const T ag = s[JC_SCAL_ALPHA_GAMMA];
    const T q = k / (s[JC_SCAL_OMH_T27] * (ag + (1.0 - ag) / (1.0 + k2 * k2)));
    const T L = jx_log(2.0 * E1 + 1.8 * q);
    const T C = 14.2 + 731.0 / (1.0 + 62.5 * q);
    return L / (L + C * q * q);
  }
  const T q = k * s[JC_SCAL_INV13KEQ];
  const T q2 = q * q;
  const T q108 = jx_exp(1.08 * (lnk - s[JC_SCAL_LN13KEQ]));
  const T c386 = 386.0 / (1.0 + 69.9 * q108);
  const T L1 = jx_log(E1 + 1.8 * s[JC_SCAL_BETA_C] * q);
  const T L2 = jx_log(E1 + 1.8 * q);
  const T C1 = 14.2 + c386, C2 = s[JC_SCAL_C14_ALPHA_C] + c386;
  const T T1 = L1 / (L1 + C1 * q2), T2 = L1 / (L1 + C2 * q2), T3 = L2 / (L2 + C1 * q2);
  const T ks = k * s[JC_SCAL_SH_D];
  const T x54 = ks / 5.4;
  const T x54_2 = x54 * x54;
  const T f = 1.0 / (1.0 + x54_2 * x54_2);
  const T Tc = f * T1 + (1.0 - f) * T2;
  const T bn = s[JC_SCAL_BETA_NODE] / ks;
  const T st = s[JC_SCAL_SH_D] * jx_rcbrt(1.0 + bn * bn * bn);
  const T x52 = ks / 5.2;
  const T bb = s[JC_SCAL_BETA_B] / ks;
  const T silk = jx_exp(-jx_exp(1.4 * (lnk - s[JC_SCAL_LNKSILK])));
  const T arg = k * st;
  const T Tb = (T3 / (1.0 + x52 * x52) + s[JC_SCAL_ALPHA_B] / (1.0 + bb * bb * bb) * silk) * (jx_sin(arg) / arg);
  return s[JC_SCAL_FB] * Tb + s[JC_SCAL_FC] * Tc;
}

template <class T> struct SetupSmem {
  T f[512];  // chi integrand at nodes+midpoints; later the normalised growth table; later (as doubles) k nodes
  T cum[256], chitab[256], gr_r[256], gr_q[256], M[127 * 4], gtab[128], sc[JC_SCAL_FIELDS];
  T d2w[JC_NHFK], S[JC_NHFR], D2[JC_NA], omm[JC_NA], ode[JC_NA], odew[JC_NA], rnl[JC_NA], red[8];
  double tab[JCM_TAB_DOUBLES];  // table-driven exp (jc_math.cuh) for the two halofit sums (~200k exp per cosmology)
  int imax[JC_NA];              // per node: number of ln k nodes inside the (k r_nl)^2 <= HF_CUT truncation
};

// =================================================================================================
// K1: per-cosmology setup.  One CTA (256 threads) per cosmology; all tables live in shared memory.
//   chi table      background.py:223-236 (RK4 on a y-independent rhs == Simpson with midpoints)
//   growth table   background.py:461-481 (RK4 == ordered product of one-step 2x2 matrices)
//   EH constants   transfer.py:47-136
//   sigma8 norm    power.py:47,56-78 (Romberg as a fixed functional)
//   halofit        power.py:86-141 (sigma^2(R,a) = D(a)^2 S(R); quirky interp root), :199-224
// =================================================================================================
#ifndef JC_SETUP_MINB
#define JC_SETUP_MINB 4  // resident CTAs per SM asked of ptxas: 64 registers (80 B spill) and 4 x 47 KB smem; 3 CTAs at 80 registers is 6 % slower
#endif
template <class T>
__global__ void __launch_bounds__(256, sizeof(T) == sizeof(double) ? JC_SETUP_MINB : (sizeof(T) <= 2 * sizeof(double) ? 2 : 1)) jc_setup_kernel(JcDevPlan pl, const double* __restrict__ cosmo,
                                                       const double* __restrict__ tangent, Ws ws, int kdiv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SetupSmem<T>& S = *reinterpret_cast<SetupSmem<T>*>(smem_raw);
  double* const s_hfk = reinterpret_cast<double*>(S.f);  // halofit k nodes (S.f is free by then)
  const int c = blockIdx.x;
  const int tid = threadIdx.x;
  const ptrdiff_t doff = ws.doff;
  for (int i = tid; i < JCM_TAB_DOUBLES; i += 256) S.tab[i] = pl.math_tab[i];  // published by the barriers below
  auto put = [&](double* p, T x) { JxMem<T>::st(p, doff, x); };
  auto node = [&](int field, int n) { return node_ptr(ws, c, field) + n; };

  // JVP passes: workspace entry c carries cosmology c / kdiv with tangent direction c % kdiv (kdiv = 1: one direction for the whole
  // pass; kdiv = K: all K directions of every cosmology in one pass).  Forward passes: kdiv = 1, tangent = nullptr.
  const double* cp = cosmo + (size_t)(c / kdiv) * pl.ncp;
  // A DualN<K> entry carries K directions: rows (c % kdiv) * K ... + K - 1 of the tangent block.
  if constexpr (JxTangents<T>::N > 0) tangent += (size_t)(c % kdiv) * JxTangents<T>::N * pl.ncp;
  T par[JC_N_COSMO_PARAMS];
#pragma unroll
  for (int i = 0; i < JC_N_COSMO_PARAMS; ++i) {
    par[i] = T(cp[i]);
    jx_seed(par[i], tangent, i, pl.ncp);
  }
  const T Oc = par[0], Ob = par[1], h = par[2], ns = par[3], s8 = par[4], Ok = par[5], w0 = par[6], wa = par[7];
  T gam = T(0.0);  // growth index (core.py:104-105), JC_GROWTH_GAMMA rows only
  if (pl.growth == JC_GROWTH_GAMMA) {
    gam = T(cp[JC_N_COSMO_PARAMS]);
    jx_seed(gam, tangent, JC_N_COSMO_PARAMS, pl.ncp);
  }
  Bg<T> bg;
  bg.Om = Ob + Oc;                 // core.py:144-146
  bg.Ok = Ok;
  bg.Ode = (1.0 - Ok) - bg.Om;     // core.py:140-150
  bg.w0 = w0; bg.wa = wa;

  // ---- EH constants: last thread, concurrently with the chi-table phase ---------------------------
  if (tid == 255) {
    const double T27 = (JC_TCMB / 2.7) * (JC_TCMB / 2.7);
    const T h2 = h * h, w_m = bg.Om * h2, w_b = Ob * h2;
    const T fb = Ob / bg.Om, fc = (bg.Om - Ob) / bg.Om;
    const T k_eq = 7.46e-2 * w_m / T27 / h;
    const T z_eq = 2.50e4 * w_m / (T27 * T27);
    const T b1 = 0.313 * jx_pow(w_m, -0.419) * (1.0 + 0.607 * jx_pow(w_m, 0.674));
    const T b2 = 0.238 * jx_pow(w_m, 0.223);
    const T z_d = 1291.0 * jx_pow(w_m, 0.251) / (1.0 + 0.659 * jx_pow(w_m, 0.828)) * (1.0 + b1 * jx_pow(w_b, b2));
    const T R_d = 31.5 * w_b / (T27 * T27) * (1.0e3 / z_d);
    const T R_eq = 31.5 * w_b / (T27 * T27) * (1.0e3 / z_eq);
    const T sh_d = 2.0 / (3.0 * k_eq) * jx_sqrt(6.0 / R_eq) *
                   jx_log((jx_sqrt(1.0 + R_d) + jx_sqrt(R_eq + R_d)) / (1.0 + jx_sqrt(R_eq)));
    const T k_silk = 1.6 * jx_pow(w_b, 0.52) * jx_pow(w_m, 0.73) * (1.0 + jx_pow(10.4 * w_m, -0.95)) / h;
    const T a1 = jx_pow(46.9 * w_m, 0.670) * (1.0 + jx_pow(32.1 * w_m, -0.532));
    const T a2 = jx_pow(12.0 * w_m, 0.424) * (1.0 + jx_pow(45.0 * w_m, -0.582));
    const T alpha_c = jx_pow(a1, -fb) * jx_pow(a2, -(fb * fb * fb));
    const T bb1 = 0.944 / (1.0 + jx_pow(458.0 * w_m, -0.708));
    const T bb2 = jx_pow(0.395 * w_m, -0.0266);
    const T beta_c = 1.0 / (1.0 + bb1 * (jx_pow(fc, bb2) - 1.0));
    const T y = (1.0 + z_eq) / (1.0 + z_d);
    const T x = jx_sqrt(1.0 + y);
    const T G = y * (-6.0 * x + (2.0 + 3.0 * y) * jx_log((x + 1.0) / (x - 1.0)));
    const T alpha_b = 2.07 * k_eq * sh_d * jx_pow(1.0 + R_d, -0.75) * G;
    const T beta_node = 8.41 * jx_pow(w_m, 0.435);
    const T beta_b = 0.5 + fb + (3.0 - 2.0 * fb) * jx_sqrt((17.2 * w_m) * (17.2 * w_m) + 1.0);
    for (int i = 0; i < JC_SCAL_FIELDS; ++i) S.sc[i] = T(0.0);
    S.sc[JC_SCAL_LN13KEQ] = jx_log(13.41 * k_eq);
    S.sc[JC_SCAL_INV13KEQ] = 1.0 / (13.41 * k_eq);
    S.sc[JC_SCAL_BETA_C] = beta_c;
    S.sc[JC_SCAL_C14_ALPHA_C] = 14.2 / alpha_c;
    S.sc[JC_SCAL_SH_D] = sh_d;
    S.sc[JC_SCAL_LNKSILK] = jx_log(k_silk);
    S.sc[JC_SCAL_ALPHA_B] = alpha_b;
    S.sc[JC_SCAL_BETA_B] = beta_b;
    S.sc[JC_SCAL_BETA_NODE] = beta_node;
    S.sc[JC_SCAL_FB] = fb;
    S.sc[JC_SCAL_FC] = fc;
    S.sc[JC_SCAL_NS] = ns;
    S.sc[JC_SCAL_OMEGA_M] = bg.Om;
    // no-wiggle fit (transfer.py:87-100): alpha_gamma and Omega_m h / (tcmb/2.7)^2
    S.sc[JC_SCAL_ALPHA_GAMMA] = 1.0 - 0.328 * jx_log(431.0 * w_m) * w_b / w_m + 0.38 * jx_log(22.3 * w_m) * (fb * fb);
    S.sc[JC_SCAL_OMH_T27] = bg.Om * h / T27;
    // the tracer kernels depend on the cosmology through (Omega_m, Omega_k, w0, wa, gamma) only: a direction without a component
    // along them has dR = 0 identically and its tangent contraction needs one product instead of two (jc_contract.cu)
    jx_flag_moves_r(S.sc[JC_SCAL_MOVES_R], tangent, pl.ncp);
  }

  // ---- chi table -------------------------------------------------------------------------------
  for (int p = tid; p < 511; p += 256) {
    const double a = pl.chi_pt_a[p], lna = pl.chi_pt_lna[p];
    T de;
    const T e2 = esqr(bg, a, lna, &de);
    S.f[p] = JC_RH / (a * a * jx_sqrt(e2)) * a;  // dchioverda(a) * a, background.py:227-229,294
  }
  __syncthreads();
  if (tid < 255) {
    const T k1 = S.f[2 * tid], k2 = S.f[2 * tid + 1], k4 = S.f[2 * tid + 2];
    S.cum[tid + 1] = pl.chi_h6[tid] * (k1 + 2.0 * k2 + 2.0 * k2 + k4);  // scipy/ode.py:19
  }
  if (tid == 0) S.cum[0] = T(0.0);
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {  // inclusive scan (np.cumsum up to re-association)
    T v = S.cum[tid];
    if (tid >= off) v = v + S.cum[tid - off];
    __syncthreads();
    S.cum[tid] = v;
    __syncthreads();
  }
  {
    const T tot = S.cum[255];
    S.chitab[tid] = tot - S.cum[tid];  // background.py:233
    put(ws.chitab + (size_t)c * JC_NCHI + tid, S.chitab[tid]);
  }

  // ---- growth table ----------------------------------------------------------------------------
  if (tid < 255) {
    const double a = pl.gr_pt_a[tid], lna = pl.gr_pt_lna[tid];
    T de;
    const T e2 = esqr(bg, a, lna, &de);
    const double ia = 1.0 / a;
    const T om = bg.Om * (ia * ia * ia) / e2;   // background.py:168
    const T ode = de / e2;                      // background.py:196
    const T w = w0 + (1.0 - a) * wa;            // background.py:52
    S.gr_q[tid] = (2.0 - 0.5 * (om + (1.0 + 3.0 * w) * ode)) / a;  // background.py:467-475
    S.gr_r[tid] = pl.growth == JC_GROWTH_GAMMA ? jx_pow(om, gam) : 1.5 * om / a / a;  // background.py:582
  }
  __syncthreads();
  if (pl.growth == JC_GROWTH_GAMMA) {
    // ln D by RK4 on a y-independent rhs f(ln a) = Omega_m(a)^gamma (background.py:538-542): per step
    // h/6 (k1 + 2 k2 + 2 k3 + k4) with k2 = k3 at the ln-a midpoint; y0 = ln a_0
    if (tid < 127) {
      const T k1 = S.gr_r[2 * tid], k2 = S.gr_r[2 * tid + 1], k4 = S.gr_r[2 * tid + 2];
      S.M[tid] = (1.0 / 6.0 * pl.gr_h[tid]) * (k1 + 2.0 * k2 + 2.0 * k2 + k4);  // scipy/ode.py:19
    }
    __syncthreads();
    if (tid == 0) {
      T y = T(pl.gr_pt_lna[0]);
      S.gtab[0] = jx_exp(y);
      for (int n = 0; n < 127; ++n) {
        y = y + S.M[n];
        S.gtab[n + 1] = jx_exp(y);
      }
    }
  } else {
  if (tid < 127) {
    const double hh = pl.gr_h[tid];
    const M2<T> A0 = {T(0.0), T(1.0), S.gr_r[2 * tid], -S.gr_q[2 * tid]};
    const M2<T> Am = {T(0.0), T(1.0), S.gr_r[2 * tid + 1], -S.gr_q[2 * tid + 1]};
    const M2<T> A1 = {T(0.0), T(1.0), S.gr_r[2 * tid + 2], -S.gr_q[2 * tid + 2]};
    const M2<T> K1 = A0;
    const M2<T> K2 = mul(Am, eye_plus(hh / 2, K1));
    const M2<T> K3 = mul(Am, eye_plus(hh / 2, K2));
    const M2<T> K4 = mul(A1, eye_plus(hh, K3));
    const double s = 1.0 / 6.0 * hh;
    S.M[4 * tid + 0] = 1.0 + s * (K1.a + 2.0 * K2.a + 2.0 * K3.a + K4.a);
    S.M[4 * tid + 1] = s * (K1.b + 2.0 * K2.b + 2.0 * K3.b + K4.b);
    S.M[4 * tid + 2] = s * (K1.c + 2.0 * K2.c + 2.0 * K3.c + K4.c);
    S.M[4 * tid + 3] = 1.0 + s * (K1.d + 2.0 * K2.d + 2.0 * K3.d + K4.d);
  }
  __syncthreads();
  // y_{n+1} = M_n ... M_0 y_0 with y_0 = (a_0, 1), background.py:477-478: inclusive prefix product of the one-step
  // matrices by a 7-step block scan (the 127 dependent matrix-vector steps of one thread held a barrier for 4 % of
  // the kernel's warp time); the product is re-associated, the table differs from the sequential scan by a few ulp
  for (int off = 1; off < 127; off <<= 1) {
    M2<T> P = {T(0.0), T(0.0), T(0.0), T(0.0)};
    if (tid < 127) {
      P = {S.M[4 * tid], S.M[4 * tid + 1], S.M[4 * tid + 2], S.M[4 * tid + 3]};
      if (tid >= off) {
        const int o = tid - off;
        const M2<T> Q = {S.M[4 * o], S.M[4 * o + 1], S.M[4 * o + 2], S.M[4 * o + 3]};
        P = mul(P, Q);  // newer steps on the left
      }
    }
    __syncthreads();
    if (tid < 127) { S.M[4 * tid] = P.a; S.M[4 * tid + 1] = P.b; S.M[4 * tid + 2] = P.c; S.M[4 * tid + 3] = P.d; }
    __syncthreads();
  }
  if (tid < 127) S.gtab[tid + 1] = S.M[4 * tid] * pl.gr_pt_a[0] + S.M[4 * tid + 1];
  if (tid == 0) S.gtab[0] = T(pl.gr_pt_a[0]);
  }  // growth ODE
  __syncthreads();
  if (tid < 128) {
    const T g = S.gtab[tid] / S.gtab[127];  // background.py:480
    put(ws.gtab + (size_t)c * JC_NGROW + tid, g);
    S.f[tid] = g;  // normalised copy (S.f is free now)
    // grid plans, ODE growth: growth-rate table f = a D'/D of the same RK4 solution (background.py:478-483)
    if (pl.grid_mode && pl.growth != JC_GROWTH_GAMMA)
      S.f[128 + tid] = tid == 0 ? T(1.0)
                                : (S.M[4 * (tid - 1) + 2] * pl.gr_pt_a[0] + S.M[4 * (tid - 1) + 3]) * pl.gr_pt_a[2 * tid] / S.gtab[tid];
  }
  __syncthreads();

  // ---- values at the 513 Limber nodes ----------------------------------------------------------
  for (int n = tid; n < JC_NA; n += 256) {
    const double a = pl.limb_a[n], lna = pl.limb_lna[n];
    int ix = pl.limb_chi_ix[n];
    const T f0 = S.chitab[ix & 255], f1 = S.chitab[ix >> 8];
    const T chi = jx_max(f0 + (f1 - f0) * pl.limb_chi_t[n], 0.0);  // background.py:242
    ix = pl.limb_gr_ix[n];
    const T g0 = S.f[ix & 255], g1 = S.f[ix >> 8];
    const T D = jx_min(jx_max(g0 + (g1 - g0) * pl.limb_gr_t[n], 0.0), 1.0);  // background.py:488
    T de;
    const T e2 = esqr(bg, a, lna, &de);
    const T se = jx_sqrt(e2);
    const T chic = pl.grid_mode ? T(1.0) : jx_max(chi, 1.0);       // angular_cl.py:73; grid plan: k is given, not (l+1/2)/chi
    const T dchida = JC_RH / (a * a * se);                         // background.py:294
    const double ia = 1.0 / a;
    const T lnchic = jx_log(chic);
    const T geom = pl.limb_w[n] * dchida / jx_max(chi * chi, 1.0) / (JC_C_LIGHT * JC_C_LIGHT);  // angular_cl.py:91,96
    put(node(JC_NODE_CHI, n), chi);
    put(node(JC_NODE_INVCHIC, n), 1.0 / chic);
    put(node(JC_NODE_LNCHIC, n), lnchic);
    if (pl.grid_mode) {  // grid plans have no Limber weight: the slot carries growth_rate(a_n) (background.py:401-440)
      const T om_a = bg.Om / (a * a * a) / e2;
      const T fa = S.f[128 + (ix & 255)], fb = S.f[128 + (ix >> 8)];
      put(node(JC_NODE_GEOM, n), pl.growth == JC_GROWTH_GAMMA ? jx_pow(om_a, gam) : fa + (fb - fa) * pl.limb_gr_t[n]);
    } else {
      put(node(JC_NODE_GEOM, n), geom);
    }
    put(node(JC_NODE_GK, n), pl.grid_mode ? T(JC_TWO_PI_SQ) : geom * JC_TWO_PI_SQ * (chic * chic * chic));  // grid plan: V = P(k, a)
    S.rnl[n] = lnchic;  // scratch until the halofit root phase
    put(node(JC_NODE_GROWTH, n), D);
    put(node(JC_NODE_HUBBLE, n), JC_H0 * se);                      // background.py:143
    S.D2[n] = D * D;
    S.omm[n] = bg.Om * (ia * ia * ia) / e2;
    S.ode[n] = de / e2;
    S.odew[n] = de / e2 * (1.0 + (w0 + (1.0 - a) * wa));
  }
  __syncthreads();  // also publishes S.sc

  // ---- sigma8 normalisation (power.py:47,70-78) -------------------------------------------------
  {
    T v = T(0.0);
    if (tid < JC_NROMB) {
      const double k = pl.romb_k[tid], lnk = pl.romb_lnk[tid];
      const T Tk = eh_transfer<T>(S.sc, k, lnk, pl.transfer);
      v = pl.romb_f[tid] * (Tk * Tk) * jx_exp(ns * lnk);
    }
    for (int o = 16; o > 0; o >>= 1) v = v + jx_shfl_xor(v, o);
    if ((tid & 31) == 0) S.red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      T s = T(0.0);
      for (int i = 0; i < 8; ++i) s = s + S.red[i];
      S.sc[JC_SCAL_SIGMASQR8] = s;
      S.sc[JC_SCAL_PKNORM] = s8 * s8 / s;
    }
    __syncthreads();
  }
  const T pknorm = S.sc[JC_SCAL_PKNORM];
  if (tid < JC_SCAL_FIELDS) put(ws.scal + (size_t)c * JC_SCAL_FIELDS + tid, S.sc[tid]);
  // separable power laws of k = (l+1/2)/chi_c: the node-side factors (the ell-side factors are plan
  // tables, and (l+1/2)^(3+n_s) goes to ws.ellpow)
  for (int n = tid; n < JC_NA; n += 256) {
    const T amp = S.D2[n] * pknorm / JC_TWO_PI_SQ;
    const T lc = S.rnl[n];
    put(node(JC_NODE_AMP, n), amp);
    put(node(JC_NODE_NQ108, n), jx_exp(-1.08 * (lc + S.sc[JC_SCAL_LN13KEQ])));
    put(node(JC_NODE_NSILK, n), jx_exp(-1.4 * (lc + S.sc[JC_SCAL_LNKSILK])));
    put(node(JC_NODE_NAMP, n), jx_exp(-(3.0 + ns) * lc) * amp);
  }
  for (int l = tid; l < pl.L; l += 256)
    put(ws.ellpow + (size_t)c * pl.Lpad + l, jx_exp((3.0 + ns) * pl.lnellp5[l]));
  if (!pl.nonlinear) return;
  __syncthreads();  // S.rnl and S.f are reused below

  // ---- halofit tables ---------------------------------------------------------------------------
  // d2w[i] = w_i * Delta^2_L(k_i, a=1) ; linear_matter_power(cosmo, k) uses growth_factor(1.0)
  const T g1sq = S.D2[JC_NA - 1];
  for (int i = tid; i < JC_NHFK; i += 256) {
    const double k = pl.hf_k[i], lnk = pl.hf_lnk[i];
    const T Tk = eh_transfer<T>(S.sc, k, lnk, pl.transfer);
    const T pk = jx_exp(ns * lnk) * (Tk * Tk) * g1sq * pknorm;  // power.py:49-52
    S.d2w[i] = pl.hf_wk[i] * (pk * (k * k * k) / JC_TWO_PI_SQ);
  }
  __syncthreads();  // every read of S.f (growth copy) is done: reuse its memory for the k nodes
  for (int i = tid; i < JC_NHFK; i += 256) s_hfk[i] = pl.hf_k[i];
  __syncthreads();
  {  // S(R_j), one R per thread (power.py:98-111 with g^2 factored out)
    // terms with (k r)^2 > HF_CUT are < e^-80 = 2e-35 of the leading ones: the k loop stops at
    // ln k <= ln sqrt(HF_CUT) - ln r (the ln k grid is uniform), which also makes it unrollable and bounds the exp
    // argument (>= -110: the unclamped table exp)
    const double r = pl.hf_r[tid];
    const double dlnk = pl.hf_lnk[1] - pl.hf_lnk[0];
    const int imax = min(JC_NHFK, (int)((HF_HALF_LN_CUT - pl.hf_logr[tid] - pl.hf_lnk[0]) / dlnk) + 2);
    T acc0 = T(0.0), acc1 = T(0.0);
    int i = 0;
    for (; i + 1 < imax; i += 2) {
      const double ya = s_hfk[i] * r, yb = s_hfk[i + 1] * r;
      acc0 = acc0 + S.d2w[i] * jcm_exp_t<false>(-(ya * ya), S.tab);
      acc1 = acc1 + S.d2w[i + 1] * jcm_exp_t<false>(-(yb * yb), S.tab);
    }
    if (i < imax) { const double ya = s_hfk[i] * r; acc0 = acc0 + S.d2w[i] * jcm_exp_t<false>(-(ya * ya), S.tab); }
    const T acc = acc0 + acc1;
    S.S[tid] = acc;
    put(ws.stab + (size_t)c * JC_NHFR + tid, acc);
  }
  __syncthreads();
  // root of sigma^2(R, a) = 1 by the reference's interp() on the DECREASING table (quirk A.9-1);
  // the index decisions use values only (frozen-index derivative)
  for (int n = tid; n < JC_NA; n += 256) {
    const T g2 = S.D2[n];
    const double g2v = jx_val(g2);
    // first index with g2*S < 1 (S decreasing); candidates for argmin((1-sig)^2) are lo-1, lo
    int lo = 0, hi = JC_NHFR;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (g2v * jx_val(S.S[mid]) >= 1.0) lo = mid + 1; else hi = mid;
    }
    int ind;
    if (lo == 0) ind = 0;
    else if (lo == JC_NHFR) ind = JC_NHFR - 1;
    else {
      const double dl = 1.0 - g2v * jx_val(S.S[lo - 1]), dr = 1.0 - g2v * jx_val(S.S[lo]);
      ind = (dl * dl <= dr * dr) ? lo - 1 : lo;  // argmin returns the first minimum
    }
    ind = min(max(ind, 1), JC_NHFR - 2);
    const T xi = g2 * S.S[ind];
    const double xq = fmin(fmax(1.0, g2v * jx_val(S.S[1])), g2v * jx_val(S.S[JC_NHFR - 2]));  // clip(x, xp[1], xp[-2])
    const int d = (xq - jx_val(xi) >= 0.0) ? 1 : -1;
    const T m = (pl.hf_logr[ind + d] - pl.hf_logr[ind]) / (g2 * S.S[ind + d] - xi);
    const T root = m * 1.0 + (pl.hf_logr[ind] - m * xi);
    const T rnl = jx_max(jx_exp(root), 1e-6);  // power.py:113-115
    S.rnl[n] = rnl;
    // truncation bound of the n_eff / C sums below, here once per node by one thread: ncu had 212 warp instructions
    // per node in the warp-per-node loop around them, most of them this bound's log() and division
    const double lnr = fmax(jx_val(root), -13.815510557964274);  // ln r_nl, ln 1e-6
    S.imax[n] = min(JC_NHFK, (int)((HF_HALF_LN_CUT - lnr - pl.hf_lnk[0]) * (256.0 / (pl.hf_lnk[JC_NHFK - 1] - pl.hf_lnk[0]))) + 2);
    put(node(JC_NODE_RNL, n), rnl);
    put(node(JC_NODE_LNKNL, n), -jx_log(rnl));
  }
  __syncthreads();
  // n_eff and C (power.py:121-141): 8 lanes per node (4 nodes per warp round), the ln k nodes strided over the 8 lanes.
  // With a whole warp per node the 5-step shuffle reduction and the store cost 60 of the ~240 warp instructions spent
  // per node (ncu region attribution, scripts/ncu_regions.py); neighbouring nodes have almost equal trip counts.
  {
    const int warp = tid >> 5, lane = tid & 31;
    auto node_sums = [&](int n, int first, int stride, T& r0, T& r1) {
      const T rnl = S.rnl[n];
      const int imax = S.imax[n];  // same (k R)^2 <= HF_CUT truncation as for S(R)
      T s1 = T(0.0), s2 = T(0.0);  // sum w e^{-y^2} y^2 and sum w e^{-y^2} y^4
#pragma unroll 2
      for (int i = first; i < imax; i += stride) {
        const T y = s_hfk[i] * rnl;
        const T y2 = y * y;
        const T t = S.d2w[i] * jx_exp_tb(-y2, S.tab) * y2;  // (k R)^2 <= ~110 by the truncation: no underflow clamp
        s1 = s1 + t;
        s2 = jx_fma(t, y2, s2);
      }
      r0 = 2.0 * s1;         // power.py:127-131: sum 2 y^2 (...)
      r1 = 4.0 * (s1 - s2);  // power.py:134-138: sum 4 (y^2 - y^4) (...)
    };
    auto finish_node = [&](int n, T r0, T r1) {
      r0 = r0 * S.D2[n]; r1 = r1 * S.D2[n];
      put(node(JC_NODE_NEFF, n), r0 - 3.0);
      put(node(JC_NODE_CURV, n), r0 * r0 + r1);
    };
    const int sub = lane & 7, grp = lane >> 3;
    for (int n0 = warp * 4; n0 < JC_NA - 1; n0 += 32) {  // nodes 0..511: 16 full rounds of 8 warps x 4 nodes
      const int n = n0 + grp;
      T r0, r1;
      node_sums(n, sub, 8, r0, r1);
      for (int o = 4; o > 0; o >>= 1) {
        r0 = r0 + jx_shfl_xor(r0, o);
        r1 = r1 + jx_shfl_xor(r1, o);
      }
      if (sub == 0) finish_node(n, r0, r1);
    }
    if (warp == 7) {  // node 512 (a = 1)
      T r0, r1;
      node_sums(JC_NA - 1, lane, 32, r0, r1);
      for (int o = 16; o > 0; o >>= 1) {
        r0 = r0 + jx_shfl_xor(r0, o);
        r1 = r1 + jx_shfl_xor(r1, o);
      }
      if (lane == 0) finish_node(JC_NA - 1, r0, r1);
    }
  }
  __syncthreads();
  // Takahashi+2012 coefficients per node (power.py:199-224, 228-238)
  const double LN10 = 2.302585092994046;
  for (int n = tid; n < JC_NA; n += 256) {
    const T ne = JxMem<T>::ld(node(JC_NODE_NEFF, n), doff), C = JxMem<T>::ld(node(JC_NODE_CURV, n), doff);
    const T n2 = ne * ne, n3 = n2 * ne, n4 = n2 * n2;
    const T odew = S.odew[n], lom = jx_log(S.omm[n]);
    if (pl.nonlinear == JC_PK_HALOFIT_SMITH2003) {  // power.py:182-198, 239-242
      const T frac = S.ode[n] / (1.0 - S.omm[n]);
      const T f1 = frac * jx_exp(-0.0307 * lom) + (1.0 - frac) * jx_exp(-0.0732 * lom);
      const T f2 = frac * jx_exp(-0.0585 * lom) + (1.0 - frac) * jx_exp(-0.1423 * lom);
      const T f3 = frac * jx_exp(0.0743 * lom) + (1.0 - frac) * jx_exp(0.0725 * lom);
      put(node(JC_NODE_AN, n), jx_exp(LN10 * (1.4861 + 1.8369 * ne + 1.6762 * n2 + 0.7940 * n3 + 0.1670 * n4 - 0.6206 * C)));
      put(node(JC_NODE_BN, n), jx_exp(LN10 * (0.9463 + 0.9466 * ne + 0.3084 * n2 - 0.9400 * C)));
      put(node(JC_NODE_LNCF, n), LN10 * (-0.2807 + 0.6669 * ne + 0.3214 * n2 - 0.0793 * C) + jx_log(f3));
      put(node(JC_NODE_P3, n), 3.0 - (0.8649 + 0.2989 * ne + 0.1631 * C));
      put(node(JC_NODE_ALPHA, n), 1.3884 + 0.3700 * ne - 0.1452 * n2);
      put(node(JC_NODE_BETA, n), 0.8291 + 0.9854 * ne + 0.3401 * n2);
      put(node(JC_NODE_MU, n), jx_exp(LN10 * (-3.5442 + 0.1908 * ne)));
      put(node(JC_NODE_NU, n), jx_exp(LN10 * (0.9585 + 1.2857 * ne)));
      put(node(JC_NODE_E1, n), 3.0 * f1);
      put(node(JC_NODE_E2, n), f2);
      continue;
    }
    put(node(JC_NODE_MU, n), T(0.0));  // power.py:223
    const T a_n = jx_exp(LN10 * (1.5222 + 2.8553 * ne + 2.3706 * n2 + 0.9903 * n3 + 0.2250 * n4 - 0.6038 * C + 0.1749 * odew));
    const T b_n = jx_exp(LN10 * (-0.5642 + 0.5864 * ne + 0.5716 * n2 - 1.5474 * C + 0.2279 * odew));
    const T lnc_n = LN10 * (0.3698 + 2.0404 * ne + 0.8161 * n2 + 0.5869 * C);
    const T gamma_n = 0.1971 - 0.0843 * ne + 0.8460 * C;
    const T alpha_n = jx_abs(6.0835 + 1.3373 * ne - 0.1959 * n2 - 5.5274 * C);
    const T beta_n = 2.0379 - 0.7354 * ne + 0.3157 * n2 + 1.2490 * n3 + 0.3980 * n4 - 0.1682 * C;
    const T nu_n = jx_exp(LN10 * (5.2105 + 3.6902 * ne));
    put(node(JC_NODE_AN, n), a_n);
    put(node(JC_NODE_BN, n), b_n);
    put(node(JC_NODE_LNCF, n), lnc_n + 0.0743 * lom);  // ln(c_n f3), f3 = om_m^0.0743
    put(node(JC_NODE_P3, n), 3.0 - gamma_n);
    put(node(JC_NODE_ALPHA, n), alpha_n);
    put(node(JC_NODE_BETA, n), beta_n);
    put(node(JC_NODE_NU, n), nu_n);
    put(node(JC_NODE_E1, n), 3.0 * jx_exp(-0.0307 * lom));  // 3 f1
    put(node(JC_NODE_E2, n), jx_exp(-0.0585 * lom));        // f2
  }
}

// T(k) of the Eisenstein-Hu fits at the plan's wavenumbers from the per-cosmology constants K1 left in ws.scal
// (grid plans: transfer.Eisenstein_Hu as a stand-alone call)
__global__ void __launch_bounds__(256) jc_transfer_kernel(JcDevPlan pl, Ws ws, double* __restrict__ tk) {
  const int c = blockIdx.y, l = blockIdx.x * 256 + threadIdx.x;
  if (l >= pl.L) return;
  tk[(size_t)c * pl.L + l] = eh_transfer<double>(ws.scal + (size_t)c * JC_SCAL_FIELDS, pl.ellp5[l], pl.lnellp5[l], pl.transfer);
}

// power.sigmasqr (power.py:56-78) for any R: one CTA per cosmology, thread = Romberg node (129 of 160), T(k) from K1's constants,
// fixed-order block reduction per R
__global__ void __launch_bounds__(160) jc_sigmasqr_kernel(JcDevPlan pl, Ws ws, const double* __restrict__ cosmo,
                                                          const double* __restrict__ R, int n_R, double* __restrict__ out) {
  __shared__ double red[5];
  const int c = blockIdx.x, tid = threadIdx.x;
  double base = 0.0, k = 1.0;
  if (tid < JC_NROMB) {
    k = pl.romb_k[tid];
    const double lnk = pl.romb_lnk[tid];
    const double Tk = eh_transfer<double>(ws.scal + (size_t)c * JC_SCAL_FIELDS, k, lnk, pl.transfer);
    base = pl.romb_w[tid] * (Tk * Tk) * exp(cosmo[(size_t)c * pl.ncp + 3] * lnk);  // pk = T^2 k^n_s (power.py:14-18, 74)
  }
  for (int r = 0; r < n_R; ++r) {
    const double x = k * R[r];
    const double w = 3.0 * (sin(x) - x * cos(x)) / (x * x * x);
    double v = base * (k * (k * w) * (k * w));
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) out[(size_t)c * n_R + r] = (((red[0] + red[1]) + (red[2] + red[3])) + red[4]) / JC_TWO_PI_SQ;
  }
}

}  // namespace

void jc_launch_sigmasqr(const JcDevPlan& pl, const Ws& ws, const double* cosmo, int chunk, const double* R_dev, int n_R, double* out,
                        cudaStream_t s) {
  jc_sigmasqr_kernel<<<chunk, 160, 0, s>>>(pl, ws, cosmo, R_dev, n_R, out);
}

void jc_launch_transfer(const JcDevPlan& pl, const Ws& ws, int chunk, double* tk, cudaStream_t s) {
  jc_transfer_kernel<<<dim3((pl.L + 255) / 256, chunk), 256, 0, s>>>(pl, ws, tk);
}

template <class T> static int setup_attr() {
  JC_CUDA_TRY(cudaFuncSetAttribute(jc_setup_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SetupSmem<T>)));
  return JC_OK;
}

int jc_setup_init() {
  static_assert(sizeof(SetupSmem<DualN<JC_JVP_MAX_GROUP>>) <= 227 * 1024, "setup tables of the widest tangent group must fit one SM");
  int st = setup_attr<DualN<1>>();
  if (st == JC_OK) st = setup_attr<DualN<2>>();
  if (st == JC_OK) st = setup_attr<DualN<3>>();
  if (st == JC_OK) st = setup_attr<DualN<4>>();
  return st;
}

void jc_launch_setup(const JcDevPlan& pl, const double* cosmo, const Ws& ws, int chunk, cudaStream_t s) {
  jc_setup_kernel<double><<<chunk, 256, sizeof(SetupSmem<double>), s>>>(pl, cosmo, nullptr, ws, 1);
}

template <class T>
static void launch_setup_t(const JcDevPlan& pl, const double* cosmo, const double* tangent, const Ws& ws, int chunk, int kdiv,
                           cudaStream_t s) {
  jc_setup_kernel<T><<<chunk, 256, sizeof(SetupSmem<T>), s>>>(pl, cosmo, tangent, ws, kdiv);
}

void jc_launch_setup_jvp(const JcDevPlan& pl, const double* cosmo, const double* tangent, const Ws& ws, int chunk, int kdiv,
                         int ntan, cudaStream_t s) {
  switch (ntan) {
    case 2: launch_setup_t<DualN<2>>(pl, cosmo, tangent, ws, chunk, kdiv, s); break;
    case 3: launch_setup_t<DualN<3>>(pl, cosmo, tangent, ws, chunk, kdiv, s); break;
    case 4: launch_setup_t<DualN<4>>(pl, cosmo, tangent, ws, chunk, kdiv, s); break;
    default: launch_setup_t<DualN<1>>(pl, cosmo, tangent, ws, chunk, kdiv, s); break;
  }
}
