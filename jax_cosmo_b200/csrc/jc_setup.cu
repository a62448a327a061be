// jc_setup.cu -- K1: per-cosmology setup kernel (background tables, EH constants, sigma8 norm, halofit).
#include "jc_internal.cuh"
#include "jc_math.cuh"

namespace {

constexpr double HF_HALF_LN_CUT = 2.1910133173369406;  // ln sqrt(80): halofit Gaussian-window truncation

struct M2 { double a, b, c, d; };  // [[a b][c d]]
__device__ __forceinline__ M2 mul(const M2& x, const M2& y) {
  return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
__device__ __forceinline__ M2 eye_plus(double s, const M2& x) {  // I + s*x
  return {1.0 + s * x.a, s * x.b, s * x.c, 1.0 + s * x.d};
}

// =================================================================================================
// K1: per-cosmology setup.  One CTA (256 threads) per cosmology; all tables live in shared memory.
//   chi table      background.py:223-236 (RK4 on a y-independent rhs == Simpson with midpoints)
//   growth table   background.py:461-481 (RK4 == ordered product of one-step 2x2 matrices)
//   EH constants   transfer.py:47-136
//   sigma8 norm    power.py:47,56-78 (Romberg as a fixed functional)
//   halofit        power.py:86-141 (sigma^2(R,a) = D(a)^2 S(R); quirky interp root), :199-224
// =================================================================================================
__global__ void __launch_bounds__(256) jc_setup_kernel(JcDevPlan pl, const double* __restrict__ cosmo,
                                                       Ws ws) {
  __shared__ double s_f[512];       // chi integrand at nodes+midpoints; reused
  __shared__ double s_cum[256];
  __shared__ double s_chitab[256];
  __shared__ double s_gr_r[256], s_gr_q[256];
  __shared__ double s_M[127 * 4];
  __shared__ double s_gtab[128];
  __shared__ double s_sc[JC_SCAL_FIELDS];
  __shared__ double s_d2w[JC_NHFK];
  __shared__ double s_S[JC_NHFR];
  __shared__ double s_D2[JC_NA];
  __shared__ double s_omm[JC_NA], s_odew[JC_NA];
  __shared__ double s_rnl[JC_NA];
  __shared__ double s_red[8];
  double* const s_hfk = s_f;  // halofit k nodes (s_f is free once the Limber-node values exist)

  const int c = blockIdx.x;
  const int tid = threadIdx.x;
  const double* cp = cosmo + (size_t)c * JC_N_COSMO_PARAMS;
  const double Oc = cp[0], Ob = cp[1], h = cp[2], ns = cp[3], s8 = cp[4], Ok = cp[5], w0 = cp[6], wa = cp[7];
  JcBg bg;
  bg.Om = Ob + Oc;                 // core.py:144-146
  bg.Ok = Ok;
  bg.Ode = (1.0 - Ok) - bg.Om;     // core.py:140-150
  bg.w0 = w0; bg.wa = wa;

  // ---- EH constants: last thread, concurrently with the chi-table phase ---------------------------
  if (tid == 255) {
    double T27 = (JC_TCMB / 2.7) * (JC_TCMB / 2.7);
    double h2 = h * h, w_m = bg.Om * h2, w_b = Ob * h2;
    double fb = Ob / bg.Om, fc = (bg.Om - Ob) / bg.Om;
    double k_eq = 7.46e-2 * w_m / T27 / h;
    double z_eq = 2.50e4 * w_m / (T27 * T27);
    double b1 = 0.313 * pow(w_m, -0.419) * (1.0 + 0.607 * pow(w_m, 0.674));
    double b2 = 0.238 * pow(w_m, 0.223);
    double z_d = 1291.0 * pow(w_m, 0.251) / (1.0 + 0.659 * pow(w_m, 0.828)) * (1.0 + b1 * pow(w_b, b2));
    double R_d = 31.5 * w_b / (T27 * T27) * (1.0e3 / z_d);
    double R_eq = 31.5 * w_b / (T27 * T27) * (1.0e3 / z_eq);
    double sh_d = 2.0 / (3.0 * k_eq) * sqrt(6.0 / R_eq) *
                  log((sqrt(1.0 + R_d) + sqrt(R_eq + R_d)) / (1.0 + sqrt(R_eq)));
    double k_silk = 1.6 * pow(w_b, 0.52) * pow(w_m, 0.73) * (1.0 + pow(10.4 * w_m, -0.95)) / h;
    double a1 = pow(46.9 * w_m, 0.670) * (1.0 + pow(32.1 * w_m, -0.532));
    double a2 = pow(12.0 * w_m, 0.424) * (1.0 + pow(45.0 * w_m, -0.582));
    double alpha_c = pow(a1, -fb) * pow(a2, -(fb * fb * fb));
    double bb1 = 0.944 / (1.0 + pow(458.0 * w_m, -0.708));
    double bb2 = pow(0.395 * w_m, -0.0266);
    double beta_c = 1.0 / (1.0 + bb1 * (pow(fc, bb2) - 1.0));
    double y = (1.0 + z_eq) / (1.0 + z_d);
    double x = sqrt(1.0 + y);
    double G = y * (-6.0 * x + (2.0 + 3.0 * y) * log((x + 1.0) / (x - 1.0)));
    double alpha_b = 2.07 * k_eq * sh_d * pow(1.0 + R_d, -0.75) * G;
    double beta_node = 8.41 * pow(w_m, 0.435);
    double beta_b = 0.5 + fb + (3.0 - 2.0 * fb) * sqrt((17.2 * w_m) * (17.2 * w_m) + 1.0);
    s_sc[JC_SCAL_LN13KEQ] = log(13.41 * k_eq);
    s_sc[JC_SCAL_INV13KEQ] = 1.0 / (13.41 * k_eq);
    s_sc[JC_SCAL_BETA_C] = beta_c;
    s_sc[JC_SCAL_C14_ALPHA_C] = 14.2 / alpha_c;
    s_sc[JC_SCAL_SH_D] = sh_d;
    s_sc[JC_SCAL_LNKSILK] = log(k_silk);
    s_sc[JC_SCAL_ALPHA_B] = alpha_b;
    s_sc[JC_SCAL_BETA_B] = beta_b;
    s_sc[JC_SCAL_BETA_NODE] = beta_node;
    s_sc[JC_SCAL_FB] = fb;
    s_sc[JC_SCAL_FC] = fc;
    s_sc[JC_SCAL_NS] = ns;
    s_sc[JC_SCAL_OMEGA_M] = bg.Om;
    for (int i = JC_SCAL_OMEGA_M + 1; i < JC_SCAL_FIELDS; ++i) s_sc[i] = 0.0;
  }

  // ---- chi table -------------------------------------------------------------------------------
  for (int p = tid; p < 511; p += 256) {
    double a = pl.chi_pt_a[p], lna = pl.chi_pt_lna[p], de;
    double e2 = jc_esqr(bg, a, lna, &de);
    s_f[p] = JC_RH / (a * a * sqrt(e2)) * a;  // dchioverda(a) * a, background.py:227-229,294
  }
  __syncthreads();
  if (tid < 255) {
    double k1 = s_f[2 * tid], k2 = s_f[2 * tid + 1], k4 = s_f[2 * tid + 2];
    s_cum[tid + 1] = pl.chi_h6[tid] * (k1 + 2 * k2 + 2 * k2 + k4);  // scipy/ode.py:19
  }
  if (tid == 0) s_cum[0] = 0.0;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {  // inclusive scan (np.cumsum up to re-association)
    double v = s_cum[tid];
    if (tid >= off) v += s_cum[tid - off];
    __syncthreads();
    s_cum[tid] = v;
    __syncthreads();
  }
  {
    double tot = s_cum[255];
    s_chitab[tid] = tot - s_cum[tid];  // background.py:233
    ws.chitab[(size_t)c * JC_NCHI + tid] = s_chitab[tid];
  }

  // ---- growth table ----------------------------------------------------------------------------
  if (tid < 255) {
    double a = pl.gr_pt_a[tid], lna = pl.gr_pt_lna[tid], de;
    double e2 = jc_esqr(bg, a, lna, &de);
    double ia = 1.0 / a;
    double om = bg.Om * (ia * ia * ia) / e2;   // background.py:168
    double ode = de / e2;                      // background.py:196
    double w = w0 + (1.0 - a) * wa;            // background.py:52
    s_gr_q[tid] = (2.0 - 0.5 * (om + (1.0 + 3.0 * w) * ode)) / a;  // background.py:467-475
    s_gr_r[tid] = 1.5 * om / a / a;
  }
  __syncthreads();
  if (tid < 127) {
    double hh = pl.gr_h[tid];
    M2 A0 = {0.0, 1.0, s_gr_r[2 * tid], -s_gr_q[2 * tid]};
    M2 Am = {0.0, 1.0, s_gr_r[2 * tid + 1], -s_gr_q[2 * tid + 1]};
    M2 A1 = {0.0, 1.0, s_gr_r[2 * tid + 2], -s_gr_q[2 * tid + 2]};
    M2 K1 = A0;
    M2 K2 = mul(Am, eye_plus(hh / 2, K1));
    M2 K3 = mul(Am, eye_plus(hh / 2, K2));
    M2 K4 = mul(A1, eye_plus(hh, K3));
    double s = 1.0 / 6.0 * hh;
    s_M[4 * tid + 0] = 1.0 + s * (K1.a + 2 * K2.a + 2 * K3.a + K4.a);
    s_M[4 * tid + 1] = s * (K1.b + 2 * K2.b + 2 * K3.b + K4.b);
    s_M[4 * tid + 2] = s * (K1.c + 2 * K2.c + 2 * K3.c + K4.c);
    s_M[4 * tid + 3] = 1.0 + s * (K1.d + 2 * K2.d + 2 * K3.d + K4.d);
  }
  __syncthreads();
  if (tid == 0) {  // ordered product applied to y0 = (a_0, 1), background.py:477-478
    double y0 = pl.gr_pt_a[0], y1 = 1.0;
    s_gtab[0] = y0;
    for (int n = 0; n < 127; ++n) {
      double n0 = s_M[4 * n] * y0 + s_M[4 * n + 1] * y1;
      double n1 = s_M[4 * n + 2] * y0 + s_M[4 * n + 3] * y1;
      y0 = n0; y1 = n1;
      s_gtab[n + 1] = y0;
    }
  }
  __syncthreads();
  if (tid < 128) {
    double g = s_gtab[tid] / s_gtab[127];  // background.py:480
    __syncwarp();
    ws.gtab[(size_t)c * JC_NGROW + tid] = g;
    s_f[tid] = g;  // normalised copy (s_f is free now)
  }
  __syncthreads();

  // ---- values at the 513 Limber nodes ----------------------------------------------------------
  for (int n = tid; n < JC_NA; n += 256) {
    double a = pl.limb_a[n], lna = pl.limb_lna[n], de;
    int ix = pl.limb_chi_ix[n];
    double f0 = s_chitab[ix & 255], f1 = s_chitab[ix >> 8];
    double chi = fmax(f0 + (f1 - f0) * pl.limb_chi_t[n], 0.0);  // background.py:242
    ix = pl.limb_gr_ix[n];
    double g0 = s_f[ix & 255], g1 = s_f[ix >> 8];
    double D = fmin(fmax(g0 + (g1 - g0) * pl.limb_gr_t[n], 0.0), 1.0);  // background.py:488
    double e2 = jc_esqr(bg, a, lna, &de);
    double se = sqrt(e2);
    double chic = fmax(chi, 1.0);                                  // angular_cl.py:73
    double dchida = JC_RH / (a * a * se);                          // background.py:294
    double ia = 1.0 / a;
    node_ptr(ws, c, JC_NODE_CHI)[n] = chi;
    node_ptr(ws, c, JC_NODE_INVCHIC)[n] = 1.0 / chic;
    const double lnchic = log(chic);
    const double geom = pl.limb_w[n] * dchida / fmax(chi * chi, 1.0) / (JC_C_LIGHT * JC_C_LIGHT);  // angular_cl.py:91,96
    node_ptr(ws, c, JC_NODE_LNCHIC)[n] = lnchic;
    node_ptr(ws, c, JC_NODE_GEOM)[n] = geom;
    node_ptr(ws, c, JC_NODE_GK)[n] = geom * JC_TWO_PI_SQ * (chic * chic * chic);
    s_rnl[n] = lnchic;  // scratch until the halofit root phase
    node_ptr(ws, c, JC_NODE_GROWTH)[n] = D;
    node_ptr(ws, c, JC_NODE_HUBBLE)[n] = JC_H0 * se;               // background.py:143
    s_D2[n] = D * D;
    s_omm[n] = bg.Om * (ia * ia * ia) / e2;
    s_odew[n] = de / e2 * (1.0 + (w0 + (1.0 - a) * wa));
  }
  __syncthreads();  // also publishes s_sc
  JcEH eh;
  jc_eh_load(eh, s_sc);

  // ---- sigma8 normalisation (power.py:47,70-78) -------------------------------------------------
  {
    double v = 0.0;
    if (tid < JC_NROMB) {
      double k = pl.romb_k[tid], lnk = pl.romb_lnk[tid];
      double Tk = jc_eh_transfer(eh, k, lnk);
      v = pl.romb_f[tid] * (Tk * Tk) * exp(ns * lnk);
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < 8; ++i) s += s_red[i];
      s_sc[JC_SCAL_SIGMASQR8] = s;
      s_sc[JC_SCAL_PKNORM] = s8 * s8 / s;
    }
    __syncthreads();
  }
  const double pknorm = s_sc[JC_SCAL_PKNORM];
  if (tid < JC_SCAL_FIELDS) ws.scal[(size_t)c * JC_SCAL_FIELDS + tid] = s_sc[tid];
  // separable power laws of k = (l+1/2)/chi_c: the node-side factors (the ell-side factors are plan
  // tables, and (l+1/2)^(3+n_s) goes to ws.ellpow)
  for (int n = tid; n < JC_NA; n += 256) {
    const double amp = s_D2[n] * pknorm / JC_TWO_PI_SQ;
    const double lc = s_rnl[n];
    node_ptr(ws, c, JC_NODE_AMP)[n] = amp;
    node_ptr(ws, c, JC_NODE_NQ108)[n] = exp(-1.08 * (lc + s_sc[JC_SCAL_LN13KEQ]));
    node_ptr(ws, c, JC_NODE_NSILK)[n] = exp(-1.4 * (lc + s_sc[JC_SCAL_LNKSILK]));
    node_ptr(ws, c, JC_NODE_NAMP)[n] = exp(-(3.0 + ns) * lc) * amp;
  }
  for (int l = tid; l < pl.L; l += 256)
    ws.ellpow[(size_t)c * pl.Lpad + l] = exp((3.0 + ns) * pl.lnellp5[l]);
  if (!pl.nonlinear) return;
  __syncthreads();  // s_rnl is reused below

  // ---- halofit tables ---------------------------------------------------------------------------
  // d2w[i] = w_i * Delta^2_L(k_i, a=1) ; linear_matter_power(cosmo, k) uses growth_factor(1.0)
  const double g1sq = s_D2[JC_NA - 1];
  for (int i = tid; i < JC_NHFK; i += 256) {
    double k = pl.hf_k[i], lnk = pl.hf_lnk[i];
    double Tk = jc_eh_transfer(eh, k, lnk);
    double pk = exp(ns * lnk) * (Tk * Tk) * g1sq * pknorm;  // power.py:49-52
    s_d2w[i] = pl.hf_wk[i] * (pk * (k * k * k) / JC_TWO_PI_SQ);
    s_hfk[i] = k;
  }
  __syncthreads();
  {  // S(R_j), one R per thread (power.py:98-111 with g^2 factored out)
    // terms with (k r)^2 > HF_CUT are < e^-80 = 2e-35 of the leading ones: the k loop stops at
    // ln k <= ln sqrt(HF_CUT) - ln r (the ln k grid is uniform), which also makes it unrollable
    const double r = pl.hf_r[tid];
    const double dlnk = pl.hf_lnk[1] - pl.hf_lnk[0];
    const int imax = min(JC_NHFK, (int)((HF_HALF_LN_CUT - pl.hf_logr[tid] - pl.hf_lnk[0]) / dlnk) + 2);
    double acc0 = 0.0, acc1 = 0.0;
    int i = 0;
    for (; i + 1 < imax; i += 2) {
      const double ya = s_hfk[i] * r, yb = s_hfk[i + 1] * r;
      acc0 = fma(s_d2w[i], jcm_exp(-(ya * ya)), acc0);
      acc1 = fma(s_d2w[i + 1], jcm_exp(-(yb * yb)), acc1);
    }
    if (i < imax) { const double ya = s_hfk[i] * r; acc0 = fma(s_d2w[i], jcm_exp(-(ya * ya)), acc0); }
    const double acc = acc0 + acc1;
    s_S[tid] = acc;
    ws.stab[(size_t)c * JC_NHFR + tid] = acc;
  }
  __syncthreads();
  // root of sigma^2(R, a) = 1 by the reference's interp() on the DECREASING table (quirk A.9-1)
  for (int n = tid; n < JC_NA; n += 256) {
    double g2 = s_D2[n];
    // first index with g2*S < 1 (S decreasing); candidates for argmin((1-sig)^2) are jj-1, jj
    int lo = 0, hi = JC_NHFR;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (g2 * s_S[mid] >= 1.0) lo = mid + 1; else hi = mid;
    }
    int ind;
    if (lo == 0) ind = 0;
    else if (lo == JC_NHFR) ind = JC_NHFR - 1;
    else {
      double dl = 1.0 - g2 * s_S[lo - 1], dr = 1.0 - g2 * s_S[lo];
      ind = (dl * dl <= dr * dr) ? lo - 1 : lo;  // argmin returns the first minimum
    }
    ind = min(max(ind, 1), JC_NHFR - 2);
    double xi = g2 * s_S[ind];
    double xq = fmin(fmax(1.0, g2 * s_S[1]), g2 * s_S[JC_NHFR - 2]);  // clip(x, xp[1], xp[-2])
    int d = (xq - xi >= 0.0) ? 1 : -1;
    double m = (pl.hf_logr[ind + d] - pl.hf_logr[ind]) / (g2 * s_S[ind + d] - xi);
    double root = m * 1.0 + (pl.hf_logr[ind] - m * xi);
    double rnl = fmax(exp(root), 1e-6);  // power.py:113-115
    s_rnl[n] = rnl;
    node_ptr(ws, c, JC_NODE_RNL)[n] = rnl;
    node_ptr(ws, c, JC_NODE_LNKNL)[n] = -log(rnl);
  }
  __syncthreads();
  // n_eff and C (power.py:121-141): one node per warp round, ln k nodes strided over lanes
  {
    const int warp = tid >> 5, lane = tid & 31;
    const double dlnk = pl.hf_lnk[1] - pl.hf_lnk[0], lnk0 = pl.hf_lnk[0];
    for (int n = warp; n < JC_NA; n += 8) {
      const double rnl = s_rnl[n];
      // same (k R)^2 <= HF_CUT truncation as for S(R); ln R_nl = -ln k_nl was stored above
      const int imax = min(JC_NHFK, (int)((HF_HALF_LN_CUT + node_ptr(ws, c, JC_NODE_LNKNL)[n] - lnk0) / dlnk) + 2);
      double r0 = 0.0, r1 = 0.0;
#pragma unroll 2
      for (int i = lane; i < imax; i += 32) {
        const double y = s_hfk[i] * rnl;
        const double y2 = y * y;
        const double res = s_d2w[i] * jcm_exp(-y2);
        r0 = fma(2.0 * res, y2, r0);
        r1 = fma(4.0 * res, y2 - y2 * y2, r1);
      }
      for (int o = 16; o > 0; o >>= 1) {
        r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
      }
      if (lane == 0) {
        r0 *= s_D2[n]; r1 *= s_D2[n];
        node_ptr(ws, c, JC_NODE_NEFF)[n] = r0 - 3.0;
        node_ptr(ws, c, JC_NODE_CURV)[n] = r0 * r0 + r1;
      }
    }
  }
  __syncthreads();
  // Takahashi+2012 coefficients per node (power.py:199-224, 228-238)
  const double LN10 = 2.302585092994046;
  for (int n = tid; n < JC_NA; n += 256) {
    double ne = node_ptr(ws, c, JC_NODE_NEFF)[n], C = node_ptr(ws, c, JC_NODE_CURV)[n];
    double n2 = ne * ne, n3 = n2 * ne, n4 = n2 * n2;
    double odew = s_odew[n], lom = log(s_omm[n]);
    double a_n = exp(LN10 * (1.5222 + 2.8553 * ne + 2.3706 * n2 + 0.9903 * n3 + 0.2250 * n4 - 0.6038 * C + 0.1749 * odew));
    double b_n = exp(LN10 * (-0.5642 + 0.5864 * ne + 0.5716 * n2 - 1.5474 * C + 0.2279 * odew));
    double lnc_n = LN10 * (0.3698 + 2.0404 * ne + 0.8161 * n2 + 0.5869 * C);
    double gamma_n = 0.1971 - 0.0843 * ne + 0.8460 * C;
    double alpha_n = fabs(6.0835 + 1.3373 * ne - 0.1959 * n2 - 5.5274 * C);
    double beta_n = 2.0379 - 0.7354 * ne + 0.3157 * n2 + 1.2490 * n3 + 0.3980 * n4 - 0.1682 * C;
    double nu_n = exp(LN10 * (5.2105 + 3.6902 * ne));
    node_ptr(ws, c, JC_NODE_AN)[n] = a_n;
    node_ptr(ws, c, JC_NODE_BN)[n] = b_n;
    node_ptr(ws, c, JC_NODE_LNCF)[n] = lnc_n + 0.0743 * lom;  // ln(c_n f3), f3 = om_m^0.0743
    node_ptr(ws, c, JC_NODE_P3)[n] = 3.0 - gamma_n;
    node_ptr(ws, c, JC_NODE_ALPHA)[n] = alpha_n;
    node_ptr(ws, c, JC_NODE_BETA)[n] = beta_n;
    node_ptr(ws, c, JC_NODE_NU)[n] = nu_n;
    node_ptr(ws, c, JC_NODE_E1)[n] = 3.0 * exp(-0.0307 * lom);  // 3 f1
    node_ptr(ws, c, JC_NODE_E2)[n] = exp(-0.0585 * lom);        // f2
  }
}

}  // namespace

void jc_launch_setup(const JcDevPlan& pl, const double* cosmo, const Ws& ws, int chunk, cudaStream_t s) {
  jc_setup_kernel<<<chunk, 256, 0, s>>>(pl, cosmo, ws);
}
