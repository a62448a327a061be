// jc_pipeline.cu -- the per-cosmology FP64 pipeline K1..K4 (see jc_internal.cuh) and its launcher.
#include <cstdio>

#include "jc_internal.cuh"

namespace {

struct Ws {  // resolved workspace pointers for one chunk
  double* chitab;  // [chunk][256]
  double* gtab;    // [chunk][128]
  double* scal;    // [chunk][32]
  double* stab;    // [chunk][256]
  double* node;    // [chunk][JC_NODE_FIELDS][JC_NA_PAD]
  double* rker;    // [chunk][T][JC_NA_PAD]
  double* vtab;    // [chunk][513][Lpad]
};

__device__ __forceinline__ double* node_ptr(const Ws& ws, int c, int field) {
  return ws.node + ((size_t)c * JC_NODE_FIELDS + field) * JC_NA_PAD;
}

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
    node_ptr(ws, c, JC_NODE_LNCHIC)[n] = log(chic);
    node_ptr(ws, c, JC_NODE_GEOM)[n] =
        pl.limb_w[n] * dchida / fmax(chi * chi, 1.0) / (JC_C_LIGHT * JC_C_LIGHT);  // angular_cl.py:91,96
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
  for (int n = tid; n < JC_NA; n += 256)
    node_ptr(ws, c, JC_NODE_AMP)[n] = s_D2[n] * pknorm / JC_TWO_PI_SQ;
  if (!pl.nonlinear) return;

  // ---- halofit tables ---------------------------------------------------------------------------
  // d2w[i] = w_i * Delta^2_L(k_i, a=1) ; linear_matter_power(cosmo, k) uses growth_factor(1.0)
  const double g1sq = s_D2[JC_NA - 1];
  for (int i = tid; i < JC_NHFK; i += 256) {
    double k = pl.hf_k[i], lnk = pl.hf_lnk[i];
    double Tk = jc_eh_transfer(eh, k, lnk);
    double pk = exp(ns * lnk) * (Tk * Tk) * g1sq * pknorm;  // power.py:49-52
    s_d2w[i] = pl.hf_wk[i] * (pk * (k * k * k) / JC_TWO_PI_SQ);
  }
  __syncthreads();
  {  // S(R_j), one R per thread (power.py:98-111 with g^2 factored out)
    double r = pl.hf_r[tid];
    double acc = 0.0;
    for (int i = 0; i < JC_NHFK; ++i) {
      double y = pl.hf_k[i] * r;
      double y2 = y * y;
      if (y2 > 300.0) break;  // exp(-300) ~ 5e-131: below any representable contribution
      acc += s_d2w[i] * exp(-y2);
    }
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
    for (int n = warp; n < JC_NA; n += 8) {
      double rnl = s_rnl[n];
      double r0 = 0.0, r1 = 0.0;
      for (int i = lane; i < JC_NHFK; i += 32) {
        double y = pl.hf_k[i] * rnl;
        double y2 = y * y;
        if (y2 > 300.0) break;
        double res = s_d2w[i] * exp(-y2);
        r0 += 2.0 * res * y2;
        r1 += 4.0 * res * (y2 - y2 * y2);
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

// =================================================================================================
// K2a: lensing efficiency  q_s(z_n) = int_{z_n}^{zmax} n_s(z') max(chi'-chi_n,0)/max(chi',1) dz'
// (probes.py:44-51) for NS sources x NCOS cosmologies per CTA; thread n owns Limber node n < 512.
// The z' grid, its chi-table brackets and n_s(z') are cosmology independent (plan tables).
// Writes the raw integral into rker[c][tracer][n]; K2b finishes the kernels.
// =================================================================================================
template <int NS, int NCOS>
__global__ void __launch_bounds__(512) jc_lens_kernel(JcDevPlan pl, Ws ws, int n_cosmo, int s0) {
  __shared__ double s_chit[NCOS][JC_NCHI];
  const int n = threadIdx.x;
  const int cb = blockIdx.x * NCOS;
  for (int i = threadIdx.x; i < NCOS * JC_NCHI; i += 512) {
    int cc = min(cb + i / JC_NCHI, n_cosmo - 1);
    s_chit[i / JC_NCHI][i % JC_NCHI] = ws.chitab[(size_t)cc * JC_NCHI + (i % JC_NCHI)];
  }
  double chin[NCOS];
#pragma unroll
  for (int c = 0; c < NCOS; ++c)
    chin[c] = node_ptr(ws, min(cb + c, n_cosmo - 1), JC_NODE_CHI)[n];
  __syncthreads();
  const size_t NL = (size_t)JC_NLENS * JC_NLENS_COLS;
  const double* nw[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) nw[s] = pl.lens_nw + (size_t)min(s0 + s, pl.n_src - 1) * NL + n;
  double acc[NCOS][NS];
#pragma unroll
  for (int c = 0; c < NCOS; ++c)
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[c][s] = 0.0;

#pragma unroll 2
  for (int m = 0; m < JC_NLENS; ++m) {
    const size_t o = (size_t)m * JC_NLENS_COLS;
    double t = pl.lens_t[o + n];
    int ix = pl.lens_ix[o + n];
    int i0 = ix & 255, i1 = ix >> 8;
    double wv[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) wv[s] = nw[s][o];
#pragma unroll
    for (int c = 0; c < NCOS; ++c) {
      double f0 = s_chit[c][i0], f1 = s_chit[c][i1];
      double chip = fmax(f0 + (f1 - f0) * t, 0.0);                   // background.py:242
      double g = fmax(chip - chin[c], 0.0) / fmax(chip, 1.0);        // probes.py:49
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[c][s] = fma(wv[s], g, acc[c][s]);
    }
  }
  const double dz = pl.lens_zmax - pl.limb_z[n];  // simps dx*N (probes.py:51)
#pragma unroll
  for (int c = 0; c < NCOS; ++c) {
    if (cb + c >= n_cosmo) break;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s0 + s >= pl.n_src) break;
      const int tr = pl.src_tracer[s0 + s];
      ws.rker[((size_t)(cb + c) * pl.T + tr) * JC_NA_PAD + n] = acc[c][s] * dz;
    }
  }
}

// K2b: finish the radial kernels R_i(a_n): WL = (q (1+z) chi 3H0^2 Om/(2c) + NLA)(1+m)
// (probes.py:51,71-74,102-129,201-207); NC = n_i(z) b_i(z) H(a) (probes.py:77-99).
__global__ void __launch_bounds__(256) jc_tracer_finish_kernel(JcDevPlan pl, Ws ws) {
  const int c = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= pl.T * JC_NA) return;
  const int t = idx / JC_NA, n = idx - t * JC_NA;
  const double Om = ws.scal[(size_t)c * JC_SCAL_FIELDS + JC_SCAL_OMEGA_M];
  const double H = node_ptr(ws, c, JC_NODE_HUBBLE)[n];
  const double D = node_ptr(ws, c, JC_NODE_GROWTH)[n];
  double* out = ws.rker + ((size_t)c * pl.T + t) * JC_NA_PAD + n;
  const double nz = pl.nz_node[(size_t)t * JC_NA_PAD + n];
  double b = pl.bias_node[(size_t)t * JC_NA_PAD + n];
  if (pl.tr_inv_growth[t]) b = b / D;  // bias.py:37-39
  double r;
  if (pl.tr_kind[t] == JC_TRACER_WEAK_LENSING) {
    const double chi = node_ptr(ws, c, JC_NODE_CHI)[n];
    const double q = (n < JC_NLENS_COLS) ? *out : 0.0;  // node 512: a=1, chi=0
    r = q * (1.0 + pl.limb_z[n]) * chi * (3.0 * JC_H0 * JC_H0 * Om / 2.0 / JC_C_LIGHT);
    if (pl.tr_ia[t]) r += nz * b * H * (-(JC_C1_RHOCRIT)*Om / D);  // probes.py:119-123
    r *= pl.tr_m1[t];
  } else {
    r = nz * b * H;
  }
  *out = r;
}

// =================================================================================================
// K3: V[n,l] = geom_n * P(k = (l+1/2)/max(chi_n,1), a_n).  One thread per (n,l), l fastest so a warp
// shares its node constants.  EH transfer (transfer.py:113-153) + linear power (power.py:49-52) +
// halofit takahashi2012 (power.py:246-262), all in registers; ln k = ln(l+1/2) - ln chi_n is shared by
// every power law.
// =================================================================================================
__global__ void __launch_bounds__(256) jc_power_kernel(JcDevPlan pl, Ws ws) {
  const int c = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= JC_NA * pl.L) return;
  const int n = idx / pl.L, l = idx - n * pl.L;
  const double* sc = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  JcEH eh;
  jc_eh_load(eh, sc);
  const double ns = sc[JC_SCAL_NS];
  const double lnk = pl.lnellp5[l] - node_ptr(ws, c, JC_NODE_LNCHIC)[n];
  const double k = pl.ellp5[l] * node_ptr(ws, c, JC_NODE_INVCHIC)[n];  // angular_cl.py:73
  const double Tk = jc_eh_transfer(eh, k, lnk);
  const double amp = node_ptr(ws, c, JC_NODE_AMP)[n];
  const double geom = node_ptr(ws, c, JC_NODE_GEOM)[n];
  const double k3 = k * k * k;
  double pk;
  if (pl.nonlinear) {
    double d2l = exp((3.0 + ns) * lnk) * (Tk * Tk) * amp;  // k^3 P_lin / (2 pi^2)
    double y = k * node_ptr(ws, c, JC_NODE_RNL)[n];
    double lny = lnk - node_ptr(ws, c, JC_NODE_LNKNL)[n];
    double beta = node_ptr(ws, c, JC_NODE_BETA)[n], alpha = node_ptr(ws, c, JC_NODE_ALPHA)[n];
    double d2q = d2l * (exp(beta * log(1.0 + d2l)) / (1.0 + alpha * d2l)) * exp(-(y / 4.0 + y * y / 8.0));
    double e1 = node_ptr(ws, c, JC_NODE_E1)[n], e2 = node_ptr(ws, c, JC_NODE_E2)[n];
    double p3 = node_ptr(ws, c, JC_NODE_P3)[n], lncf = node_ptr(ws, c, JC_NODE_LNCF)[n];
    double d2hp = node_ptr(ws, c, JC_NODE_AN)[n] * exp(e1 * lny) /
                  (1.0 + node_ptr(ws, c, JC_NODE_BN)[n] * exp(e2 * lny) + exp(p3 * (lncf + lny)));
    double d2h = d2hp / (1.0 + node_ptr(ws, c, JC_NODE_NU)[n] / (y * y));
    pk = JC_TWO_PI_SQ / k3 * (d2q + d2h);  // power.py:260
  } else {
    pk = exp(ns * lnk) * (Tk * Tk) * (amp * JC_TWO_PI_SQ);  // power.py:49-52
  }
  ws.vtab[((size_t)c * JC_NA + n) * pl.Lpad + l] = pk * geom;
}

// =================================================================================================
// K4: pair contraction  C[(i,j),l] = e_i(l) e_j(l) sum_n R_i[n] R_j[n] V[n,l]   (angular_cl.py:82-96)
// CTA = (cosmology, tile of 32 ell).  R (all tracers) and the V tile are staged in shared memory;
// a warp owns one 4x4 block of tracer pairs, lanes = ell, so R loads are warp-broadcast LDS.128 and
// V loads are conflict-free.  16 accumulators per thread, 20 FP64 ops per 5 LDS.
// =================================================================================================
#define JC_CT_WARPS 16
__global__ void __launch_bounds__(JC_CT_WARPS * 32) jc_contract_kernel(JcDevPlan pl, Ws ws, double* __restrict__ cl,
                                                                      int nch) {
  extern __shared__ __align__(16) double smem[];
  const int c = blockIdx.y;
  const int l0 = blockIdx.x * 32;
  const int TP = (pl.T + 3) & ~3;
  const int TS = TP + 2;  // row stride: 16-byte aligned rows, <=2-way bank conflicts on the fill
  const int nb = TP >> 2;
  const int ntask = nb * (nb + 1) / 2;
  double* Rs = smem;                     // [nch][TS]
  double* Vs = smem + (size_t)nch * TS;  // [nch][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = l0 + lane;
  const bool lok = l < pl.L;
  const double* Rg = ws.rker + (size_t)c * pl.T * JC_NA_PAD;
  const double* Vg = ws.vtab + (size_t)c * JC_NA * pl.Lpad;

  for (int round = 0; round * JC_CT_WARPS < ntask; ++round) {
    const int task = round * JC_CT_WARPS + warp;
    int bi = 0, bj = 0;
    if (task < ntask) {  // task -> (bi <= bj), row-major upper triangle of blocks
      int t = task;
      while (t >= nb - bi) { t -= nb - bi; ++bi; }
      bj = bi + t;
    }
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int n0 = 0; n0 < JC_NA; n0 += nch) {
      const int nn = min(nch, JC_NA - n0);
      __syncthreads();
      for (int i = threadIdx.x; i < nn * TP; i += blockDim.x) {
        int t = i / nn, n = i - t * nn;  // n fastest: coalesced global reads
        Rs[n * TS + t] = (t < pl.T) ? Rg[(size_t)t * JC_NA_PAD + n0 + n] : 0.0;
      }
      for (int i = threadIdx.x; i < nn * 32; i += blockDim.x) {
        int n = i >> 5, ll = i & 31;
        Vs[i] = (l0 + ll < pl.L) ? Vg[(size_t)(n0 + n) * pl.Lpad + l0 + ll] : 0.0;
      }
      __syncthreads();
      if (task < ntask) {
        const double* ri = Rs + 4 * bi;
        const double* rj = Rs + 4 * bj;
#pragma unroll 4
        for (int n = 0; n < nn; ++n) {
          double v = Vs[n * 32 + lane];
          double2 a01 = *reinterpret_cast<const double2*>(ri + n * TS);
          double2 a23 = *reinterpret_cast<const double2*>(ri + n * TS + 2);
          double2 b01 = *reinterpret_cast<const double2*>(rj + n * TS);
          double2 b23 = *reinterpret_cast<const double2*>(rj + n * TS + 2);
          double a[4] = {a01.x, a01.y, a23.x, a23.y};
          double b[4] = {b01.x * v, b01.y * v, b23.x * v, b23.y * v};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
      }
    }
    if (task < ntask && lok) {
      const double ef = pl.ellfac[l];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ti = 4 * bi + i;
        if (ti >= pl.T) break;
        const double ei = pl.tr_kind[ti] == JC_TRACER_WEAK_LENSING ? ef : 1.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tj = 4 * bj + j;
          if (tj >= pl.T || tj < ti) continue;
          const double ej = pl.tr_kind[tj] == JC_TRACER_WEAK_LENSING ? ef : 1.0;
          const int p = ti * pl.T - (ti * (ti - 1)) / 2 + (tj - ti);
          cl[((size_t)c * pl.P + p) * pl.L + l] = acc[i][j] * (ei * ej);
        }
      }
    }
  }
}

template <int NS>
void launch_lens(const JcDevPlan& pl, const Ws& ws, int chunk, int s0, cudaStream_t st) {
  constexpr int NCOS = 2;
  jc_lens_kernel<NS, NCOS><<<(chunk + NCOS - 1) / NCOS, 512, 0, st>>>(pl, ws, chunk, s0);
}

size_t per_cosmo_doubles(const JcDevPlan& pl) {
  return (size_t)JC_NCHI + JC_NGROW + JC_SCAL_FIELDS + JC_NHFR + (size_t)JC_NODE_FIELDS * JC_NA_PAD +
         (size_t)pl.T * JC_NA_PAD + (size_t)JC_NA * pl.Lpad;
}

void layout_for(const JcDevPlan& pl, int64_t chunk, jc_ws_layout* lo) {
  int64_t o = 0;
  lo->chunk = chunk;
  lo->node_stride = JC_NA_PAD;
  lo->ell_stride = pl.Lpad;
  lo->chitab = o; o += chunk * JC_NCHI;
  lo->gtab = o; o += chunk * JC_NGROW;
  lo->scal = o; o += chunk * JC_SCAL_FIELDS;
  lo->stab = o; o += chunk * JC_NHFR;
  lo->node = o; o += chunk * (int64_t)JC_NODE_FIELDS * JC_NA_PAD;
  lo->rker = o; o += chunk * (int64_t)pl.T * JC_NA_PAD;
  lo->vtab = o; o += chunk * (int64_t)JC_NA * pl.Lpad;
  lo->total = o;
}

}  // namespace

int jc_pipeline_init() {
  JC_CUDA_TRY(cudaFuncSetAttribute(jc_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return JC_OK;
}

extern "C" int jc_workspace_bytes(const jc_plan* plan, int64_t n_cosmo, size_t* bytes_out) {
  if (!plan || !bytes_out || n_cosmo < 1) return JC_ERR_INVALID;
  int64_t chunk = n_cosmo < JC_MAX_CHUNK ? n_cosmo : JC_MAX_CHUNK;
  *bytes_out = per_cosmo_doubles(plan->d) * (size_t)chunk * sizeof(double);
  return JC_OK;
}

extern "C" int jc_workspace_layout(const jc_plan* plan, size_t ws_bytes, jc_ws_layout* lo) {
  if (!plan || !lo) return JC_ERR_INVALID;
  size_t per = per_cosmo_doubles(plan->d) * sizeof(double);
  int64_t chunk = (int64_t)(ws_bytes / per);
  if (chunk < 1) return JC_ERR_WORKSPACE;
  if (chunk > JC_MAX_CHUNK) chunk = JC_MAX_CHUNK;
  layout_for(plan->d, chunk, lo);
  return JC_OK;
}

extern "C" int jc_angular_cl_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo,
                                 double* cl_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!plan || !cosmo_dev || !cl_dev || !ws_dev || n_cosmo < 1) return JC_ERR_INVALID;
  jc_ws_layout lo;
  int st = jc_workspace_layout(plan, ws_bytes, &lo);
  if (st != JC_OK) return st;
  const JcDevPlan& pl = plan->d;
  cudaStream_t s = (cudaStream_t)stream;
  double* base = (double*)ws_dev;
  Ws ws;
  ws.chitab = base + lo.chitab; ws.gtab = base + lo.gtab; ws.scal = base + lo.scal;
  ws.stab = base + lo.stab; ws.node = base + lo.node; ws.rker = base + lo.rker; ws.vtab = base + lo.vtab;

  // contraction kernel shared-memory plan
  const int TS = ((pl.T + 3) & ~3) + 2;
  int nch = (int)((220 * 1024) / (sizeof(double) * (TS + 32)));
  if (nch > JC_NA) nch = JC_NA;
  size_t ct_smem = (size_t)nch * (TS + 32) * sizeof(double);

  JcProf* prof = (plan->prof && plan->prof->enabled) ? plan->prof : nullptr;
  for (int64_t c0 = 0; c0 < n_cosmo; c0 += lo.chunk) {
    const int chunk = (int)((n_cosmo - c0) < lo.chunk ? (n_cosmo - c0) : lo.chunk);
    const double* cos = cosmo_dev + c0 * JC_N_COSMO_PARAMS;
    cudaEvent_t* ev = nullptr;
    int* nl = nullptr;
    if (prof && prof->used < JC_PROF_SLOTS) { ev = prof->ev[prof->used]; nl = prof->launches[prof->used]; ++prof->used; }
#define JC_MARK(i) do { if (ev) cudaEventRecord(ev[i], s); } while (0)
    JC_MARK(0);
    jc_setup_kernel<<<chunk, 256, 0, s>>>(pl, cos, ws);
    JC_MARK(1);
    int n_lens = 0;
    for (int s0 = 0; s0 < pl.n_src; ++n_lens) {
      int rem = pl.n_src - s0;
      if (rem >= 5) { launch_lens<5>(pl, ws, chunk, s0, s); s0 += 5; }
      else if (rem == 4) { launch_lens<4>(pl, ws, chunk, s0, s); s0 += 4; }
      else if (rem == 3) { launch_lens<3>(pl, ws, chunk, s0, s); s0 += 3; }
      else if (rem == 2) { launch_lens<2>(pl, ws, chunk, s0, s); s0 += 2; }
      else { launch_lens<1>(pl, ws, chunk, s0, s); s0 += 1; }
    }
    JC_MARK(2);
    jc_tracer_finish_kernel<<<dim3((pl.T * JC_NA + 255) / 256, chunk), 256, 0, s>>>(pl, ws);
    JC_MARK(3);
    jc_power_kernel<<<dim3((JC_NA * pl.L + 255) / 256, chunk), 256, 0, s>>>(pl, ws);
    JC_MARK(4);
    jc_contract_kernel<<<dim3((pl.L + 31) / 32, chunk), JC_CT_WARPS * 32, ct_smem, s>>>(
        pl, ws, cl_dev + (size_t)c0 * pl.P * pl.L, nch);
    JC_MARK(5);
#undef JC_MARK
    if (nl) { nl[0] = 1; nl[1] = n_lens; nl[2] = 1; nl[3] = 1; nl[4] = 1; }
  }
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}
