// jc_plan.cu -- plan creation: validation of the problem, cosmology-independent quadrature grids
// and interpolation brackets (host, C++), one-time n(z) kernels (device).
//
// The grids restate the reference's discretisation exactly:
//   chi table    logspace(-3,0,256) in a, RK4 in ln a            background.py:223-236
//   growth table logspace(-3,0,128), RK4 in a                    background.py:461-481
//   Limber       simps over linspace(1/(1+zmax), 1, 513)         angular_cl.py:96
//   lensing      simps over linspace(z_n, zmax, 257) per node    probes.py:51
//   sigma8       Romberg divmax=7 over x in [-4,3], k = e^x      power.py:70-78
//   halofit      ln k in linspace(ln1e-4, ln1e4, 257), ln R in linspace(ln1e-4, ln10, 256)
//                                                                power.py:93,111
//   interp       nearest node + neighbour by sign                scipy/interpolate.py:25-37
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include <algorithm>

#include "jc_internal.cuh"
#include "jc_math.cuh"

namespace {

// numpy.linspace(start, stop, num) for float64 (i*step + start, last = stop)
std::vector<double> linspace(double start, double stop, int num) {
  std::vector<double> y(num);
  int div = num - 1;
  double delta = stop - start;
  double step = delta / div;
  for (int i = 0; i < num; ++i) {
    volatile double p = (double)i * step;  // no FMA contraction
    y[i] = p + start;
  }
  if (num > 1) y[num - 1] = stop;
  return y;
}

// Composite Simpson weights (1,4,2,...,4,1) * dx/3, scipy/integrate.py:193-199
std::vector<double> simpson_weights(int N, double dx) {
  std::vector<double> w(N + 1, 1.0);
  for (int i = 1; i < N; ++i) w[i] = (i & 1) ? 4.0 : 2.0;
  for (auto& v : w) v *= dx / 3.0;
  return w;
}

// Romberg (scipy/integrate.py:134-159) as weights over 2^divmax+1 equispaced nodes, unit interval
std::vector<double> romberg_weights(int divmax) {
  int n = (1 << divmax) + 1;
  std::vector<std::vector<double>> R;
  for (int i = 0; i <= divmax; ++i) {
    int step = 1 << (divmax - i);
    std::vector<double> w(n, 0.0);
    for (int j = 0; j < n; j += step) w[j] = 1.0;
    w[0] = w[n - 1] = 0.5;
    for (auto& v : w) v /= (double)(1 << i);
    R.push_back(w);
  }
  for (int k = 1; k <= divmax; ++k) {
    double f = std::pow(4.0, k);
    std::vector<std::vector<double>> Rn;
    for (size_t j = 0; j + 1 < R.size(); ++j) {
      std::vector<double> w(n);
      for (int t = 0; t < n; ++t) w[t] = (f * R[j + 1][t] - R[j][t]) / (f - 1.0);
      Rn.push_back(w);
    }
    R.swap(Rn);
  }
  return R[0];
}

// The (ind, ind+d) pair chosen by scipy/interpolate.py:25-37 for an increasing table, and the
// weight t such that interp(x) = fp[i0] + (fp[i1]-fp[i0]) * t.
void bracket(double x, const std::vector<double>& xp, int* i0, int* i1, double* t) {
  int n = (int)xp.size();
  // searchsorted (left)
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (xp[mid] < x) lo = mid + 1; else hi = mid;
  }
  int j = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);
  double dl = (x - xp[j - 1]) * (x - xp[j - 1]);
  double dr = (x - xp[j]) * (x - xp[j]);
  int ind = (dl <= dr) ? j - 1 : j;  // argmin returns the first minimum
  if (ind < 1) ind = 1;
  if (ind > n - 2) ind = n - 2;
  double xc = x < xp[1] ? xp[1] : (x > xp[n - 2] ? xp[n - 2] : x);
  int d = (xc - xp[ind] >= 0.0) ? 1 : -1;
  *i0 = ind;
  *i1 = ind + d;
  *t = (x - xp[ind]) / (xp[ind + d] - xp[ind]);
}

struct Blob {
  std::vector<unsigned char> host;
  size_t add(const void* p, size_t bytes) {
    size_t off = (host.size() + 255) & ~(size_t)255;
    host.resize(off + bytes);
    if (p) memcpy(host.data() + off, p, bytes);
    return off;
  }
  template <class T>
  size_t add(const std::vector<T>& v) { return add(v.data(), v.size() * sizeof(T)); }
  size_t reserve(size_t bytes) { return add(nullptr, bytes); }
};

// ---- one-time device kernels ----------------------------------------------------------------
struct NzDev {
  int family, n_shifts;
  double p[4];
  double shifts[JC_MAX_SHIFTS];
  double zmax;
  const double* kde_z;  // device copies of the catalogue (JC_NZ_KDE)
  const double* kde_w;
  long long kde_n;
  double kde_bw;
};
struct NzDevAll { NzDev nz[JC_MAX_TRACERS]; };

__device__ __forceinline__ double pz_fn(const NzDev& nz, double z) {
  // systematic_shift chain (redshift.py:169-171), then the family's un-normalised n(z)
  for (int s = 0; s < nz.n_shifts; ++s) z = fmax(z - nz.shifts[s], 0.0);
  switch (nz.family) {
    case JC_NZ_FU:  // redshift.py:103-105
      return (pow(z, nz.p[0]) + pow(z, nz.p[0] * nz.p[1])) / (pow(z, nz.p[1]) + nz.p[2]);
    case JC_NZ_KDE: {  // redshift.py:142-156
      const double bw = nz.kde_bw, norm = 1.0 / sqrt(2.0 * 3.141592653589793) / bw;
      double s = 0.0, q = 0.0;
      for (long long i = 0; i < nz.kde_n; ++i) {
        const double d = nz.kde_z[i] - z;
        s += nz.kde_w[i] * (norm * exp(-(d * d) / (bw * bw * 2.0)));
        q += nz.kde_w[i];
      }
      return s / q;
    }
    case JC_NZ_DELTA:  // never evaluated as a distribution (lensing uses the source plane directly)
      return 0.0;
    default:  // JC_NZ_SMAIL, redshift.py:75-77
      return pow(z, nz.p[0]) * exp(-pow(z / nz.p[2], nz.p[1]));
  }
}

// norm[t] = simps(pz_fn, 0, zmax, 256)  (redshift.py:29-30); one block of 256+ threads per tracer
__global__ void jc_nz_norm_kernel(NzDevAll all, double* __restrict__ norm) {
  __shared__ double red[288];
  const NzDev& nz = all.nz[blockIdx.x];
  int i = threadIdx.x;
  double v = 0.0;
  if (i <= 256) {
    double dx = nz.zmax / 256.0;
    double z = (i == 256) ? nz.zmax : (double)i * dx;  // linspace(0, zmax, 257)
    double w = (i == 0 || i == 256) ? 1.0 : ((i & 1) ? 4.0 : 2.0);
    v = w * pz_fn(nz, z);
  }
  red[i] = v;
  __syncthreads();
  if (i == 0) {
    double s = 0.0;
    for (int j = 0; j <= 256; ++j) s += red[j];
    norm[blockIdx.x] = nz.family == JC_NZ_DELTA ? 1.0 : nz.zmax / 256.0 / 3.0 * s;  // redshift.py:118
  }
}

// nz_node[n][t] = pz_t(z_n)/norm_t on the 513 Limber nodes (node-major, stride TS, like ws.rker)
__global__ void jc_nz_node_kernel(NzDevAll all, const double* __restrict__ norm,
                                  const double* __restrict__ limb_z, double* __restrict__ out, int TS) {
  int t = blockIdx.y;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < JC_NA) out[(size_t)n * TS + t] = pz_fn(all.nz[t], limb_z[n]) / norm[t];
}

// lens_nw[s][m][n] = simpson_w[m]/(3*256) * pz_s(z'(m,n))/norm_s ; z' = linspace(z_n, zmax, 257)[m]
__global__ void jc_nz_lens_kernel(NzDevAll all, const int* __restrict__ src_tracer,
                                  const double* __restrict__ norm,
                                  const double* __restrict__ lens_z, double* __restrict__ out) {
  int s = blockIdx.z;
  int m = blockIdx.y;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= JC_NLENS_COLS) return;
  int t = src_tracer[s];
  double w = (m == 0 || m == 256) ? 1.0 : ((m & 1) ? 4.0 : 2.0);
  size_t o = (size_t)m * JC_NLENS_COLS + n;
  out[(size_t)s * JC_NLENS * JC_NLENS_COLS + o] =
      (w / 3.0 / 256.0) * (pz_fn(all.nz[t], lens_z[o]) / norm[t]);
}

int validate(const jc_problem* pb, int n_ell) {
  if (!pb) return JC_ERR_INVALID;
  if (pb->abi_version != JC_ABI_VERSION) return JC_ERR_INVALID;
  if (pb->n_tracers < 1 || pb->n_tracers > JC_MAX_TRACERS) return JC_ERR_INVALID;
  if (n_ell < 1) return JC_ERR_INVALID;
  if (pb->transfer != JC_TF_EISENSTEIN_HU_OSC && pb->transfer != JC_TF_EISENSTEIN_HU_NOWIGGLE) return JC_ERR_UNSUPPORTED;
  if (pb->nonlinear < JC_PK_LINEAR || pb->nonlinear > JC_PK_HALOFIT_SMITH2003) return JC_ERR_UNSUPPORTED;
  if (pb->growth != JC_GROWTH_ODE && pb->growth != JC_GROWTH_GAMMA) return JC_ERR_UNSUPPORTED;
  double lens_zmax = -1.0;
  for (int t = 0; t < pb->n_tracers; ++t) {
    const jc_tracer& tr = pb->tracers[t];
    if (tr.kind != JC_TRACER_WEAK_LENSING && tr.kind != JC_TRACER_NUMBER_COUNTS)
      return JC_ERR_INVALID;
    if (tr.nz.family < JC_NZ_SMAIL || tr.nz.family > JC_NZ_KDE) return JC_ERR_UNSUPPORTED;
    if (tr.nz.family == JC_NZ_KDE && (!tr.nz.kde_z || !tr.nz.kde_w || tr.nz.kde_n < 1 || !(tr.nz.kde_bw > 0.0)))
      return JC_ERR_INVALID;
    // delta planes: weak lensing without IA only, not under a shift (the reference raises NotImplementedError
    // in density_kernel / nla_kernel, probes.py:82-85,107-110)
    if (tr.nz.family == JC_NZ_DELTA &&
        (tr.kind != JC_TRACER_WEAK_LENSING || tr.ia_enabled || tr.nz.n_shifts != 0 || !(tr.nz.params[0] >= 0.0)))
      return JC_ERR_UNSUPPORTED;
    if (tr.nz.n_shifts < 0 || tr.nz.n_shifts > JC_MAX_SHIFTS) return JC_ERR_UNSUPPORTED;
    if (!(tr.nz.zmax > 0.0) || !(tr.probe_zmax > 0.0)) return JC_ERR_INVALID;
    if (!(tr.nz.gals_per_arcmin2 > 0.0)) return JC_ERR_INVALID;
    bool needs_bias = tr.kind == JC_TRACER_NUMBER_COUNTS || tr.ia_enabled;
    if (needs_bias) {
      if (tr.bias.family < JC_BIAS_CONSTANT || tr.bias.family > JC_BIAS_DES_Y1_IA)
        return JC_ERR_INVALID;
    }
    if (tr.kind == JC_TRACER_WEAK_LENSING) {
      // one z' grid per plan: all WL probes must share their zmax (true for the default zmax=10)
      if (lens_zmax < 0.0) lens_zmax = tr.probe_zmax;
      else if (lens_zmax != tr.probe_zmax) return JC_ERR_UNSUPPORTED;
    }
  }
  return JC_OK;
}

}  // namespace

void jc_math_table(double* out) {
  for (int j = 0; j < JCM_EXP_N; ++j) out[JCM_TAB_EXP + j] = std::exp2(j / (double)JCM_EXP_N);
  for (int j = 0; j < 128; ++j) {
    const double cj = j == 0 ? 1.0 : 1.0 / (1.0 + (j + 0.5) / 128.0);
    out[JCM_TAB_LOG + 2 * j] = cj;
    out[JCM_TAB_LOG + 2 * j + 1] = j == 0 ? 0.0 : -std::log(cj);
  }
}

// grid_a != nullptr: grid plan -- the 513 "Limber nodes" are the caller's n_grid_a <= 512 scale factors (padded with
// a = 1, which node 512 must be: halofit normalises sigma^2(R) with D(1)) and ell_host holds wavenumbers k.
static int create_plan(const jc_problem* pb, const double* ell_host, int32_t n_ell, int32_t device,
                       const double* grid_a, int32_t n_grid_a, jc_plan** plan_out) {
  if (!plan_out || !ell_host) return JC_ERR_INVALID;
  *plan_out = nullptr;
  int st = validate(pb, n_ell);
  if (st != JC_OK) return st;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return JC_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return JC_ERR_INVALID;
  JcDeviceGuard guard(device);
  JC_CUDA_TRY(guard.status);
  st = jc_pipeline_init();
  if (st != JC_OK) return st;

  const int T = pb->n_tracers, L = n_ell;
  const int P = T * (T + 1) / 2;
  double zmax = 0.0, lens_zmax = 0.0;
  int n_src = 0;
  std::vector<int> tr_kind(T), tr_inv(T, 0), tr_ia(T, 0), tr_src(T, -1), src_tracer, tr_delta_ix(T, -1);
  std::vector<double> tr_m1(T, 1.0), tr_delta_t(T, 0.0);
  for (int t = 0; t < T; ++t) {
    const jc_tracer& tr = pb->tracers[t];
    if (tr.probe_zmax > zmax) zmax = tr.probe_zmax;  // angular_cl.py:63
    tr_kind[t] = tr.kind;
    bool needs_bias = tr.kind == JC_TRACER_NUMBER_COUNTS || tr.ia_enabled;
    tr_inv[t] = needs_bias && tr.bias.family == JC_BIAS_INVERSE_GROWTH;
    if (tr.kind == JC_TRACER_WEAK_LENSING) {
      tr_ia[t] = tr.ia_enabled ? 1 : 0;
      if (tr.nz.family != JC_NZ_DELTA) {  // extended distributions go through the lensing-efficiency integral
        tr_src[t] = n_src++;
        src_tracer.push_back(t);
      }
      tr_m1[t] = 1.0 + tr.m_bias;
      lens_zmax = tr.probe_zmax;
    }
  }
  if (src_tracer.empty()) src_tracer.push_back(0);

  Blob B;
  JcDevPlan d;
  memset(&d, 0, sizeof(d));
  d.T = T; d.P = P; d.L = L; d.Lpad = (L + 3) & ~3; d.nonlinear = pb->nonlinear; d.transfer = pb->transfer;
  d.growth = pb->growth; d.ncp = JC_N_COSMO_PARAMS + (pb->growth == JC_GROWTH_GAMMA ? 1 : 0);
  d.grid_mode = grid_a ? 1 : 0; d.grid_na = grid_a ? n_grid_a : 0;
  d.TS = T;  // bank-conflict-free A-fragment gathers in the contraction kernel need TS = 4 or 12 (mod 16)
  while (d.TS % 16 != 4 && d.TS % 16 != 12) ++d.TS;
  d.n_src = n_src; d.zmax = zmax; d.lens_zmax = lens_zmax;

  // ---- chi table grid -----------------------------------------------------------------------
  std::vector<double> e256 = linspace(-3.0, 0.0, JC_NCHI), atab(JC_NCHI), xtab(JC_NCHI);
  for (int i = 0; i < JC_NCHI; ++i) { atab[i] = std::pow(10.0, e256[i]); xtab[i] = std::log(atab[i]); }
  for (int t = 0; t < T; ++t) {  // delta_nz source planes: chi-table bracket of a_s = 1/(1+z_s) (probes.py:57-59)
    const jc_tracer& tr = pb->tracers[t];
    if (tr.kind == JC_TRACER_WEAK_LENSING && tr.nz.family == JC_NZ_DELTA) {
      int i0, i1;
      bracket(1.0 / (1.0 + tr.nz.params[0]), atab, &i0, &i1, &tr_delta_t[t]);
      tr_delta_ix[t] = i0 | (i1 << 8);
    }
  }
  std::vector<double> chi_pt_a(511), chi_pt_lna(511), chi_h6(255);
  for (int i = 0; i < JC_NCHI; ++i) { chi_pt_a[2 * i] = atab[i]; chi_pt_lna[2 * i] = xtab[i]; }
  for (int i = 0; i < JC_NCHI - 1; ++i) {
    double h = xtab[i + 1] - xtab[i];
    double xm = xtab[i] + h / 2;
    chi_pt_lna[2 * i + 1] = xm;
    chi_pt_a[2 * i + 1] = std::exp(xm);
    chi_h6[i] = 1.0 / 6.0 * h;
  }
  // ---- growth table grid ----------------------------------------------------------------------
  std::vector<double> e128 = linspace(-3.0, 0.0, JC_NGROW), ag(JC_NGROW);
  for (int i = 0; i < JC_NGROW; ++i) ag[i] = std::pow(10.0, e128[i]);
  std::vector<double> gr_pt_a(255), gr_pt_lna(255), gr_h(127);
  for (int i = 0; i < JC_NGROW; ++i) gr_pt_a[2 * i] = ag[i];
  if (pb->growth == JC_GROWTH_GAMMA) {
    // odeint over t = log(atab) (background.py:536-542): steps and midpoints in ln a, xa = exp(loga)
    for (int i = 0; i < JC_NGROW; ++i) { gr_pt_lna[2 * i] = std::log(ag[i]); gr_pt_a[2 * i] = std::exp(gr_pt_lna[2 * i]); }
    for (int i = 0; i < JC_NGROW - 1; ++i) {
      gr_h[i] = gr_pt_lna[2 * i + 2] - gr_pt_lna[2 * i];
      gr_pt_lna[2 * i + 1] = gr_pt_lna[2 * i] + gr_h[i] / 2;
      gr_pt_a[2 * i + 1] = std::exp(gr_pt_lna[2 * i + 1]);
    }
  } else {
    for (int i = 0; i < JC_NGROW - 1; ++i) {
      gr_h[i] = ag[i + 1] - ag[i];
      gr_pt_a[2 * i + 1] = ag[i] + gr_h[i] / 2;
    }
    for (int i = 0; i < 255; ++i) gr_pt_lna[i] = std::log(gr_pt_a[i]);
  }
  // ---- Limber nodes ---------------------------------------------------------------------------
  double amin = 1.0 / (1.0 + zmax);  // z2a(zmax), utils.py:2-4
  std::vector<double> la = linspace(amin, 1.0, JC_NA), llna(JC_NA), lz(JC_NA);
  if (grid_a)
    for (int n = 0; n < JC_NA; ++n) la[n] = n < n_grid_a ? grid_a[n] : 1.0;
  std::vector<double> lw = simpson_weights(512, (1.0 - amin) / 512);
  std::vector<double> lct(JC_NA), lgt(JC_NA);
  std::vector<uint16_t> lcix(JC_NA), lgix(JC_NA);
  for (int n = 0; n < JC_NA; ++n) {
    llna[n] = std::log(la[n]);
    lz[n] = 1.0 / la[n] - 1.0;  // a2z, utils.py:7-9
    int i0, i1; double t;
    bracket(la[n], atab, &i0, &i1, &t);
    lcix[n] = (uint16_t)(i0 | (i1 << 8)); lct[n] = t;
    bracket(la[n], ag, &i0, &i1, &t);
    lgix[n] = (uint16_t)(i0 | (i1 << 8)); lgt[n] = t;
  }
  // ---- sigma8 Romberg functional ----------------------------------------------------------------
  std::vector<double> rw = romberg_weights(7), rk(JC_NROMB), rlnk(JC_NROMB), rf(JC_NROMB), rwn(JC_NROMB);
  {
    double lo = std::log10(0.0001), hi = std::log10(1000.0);
    std::vector<double> x = linspace(lo, hi, JC_NROMB);
    for (int i = 0; i < JC_NROMB; ++i) {
      double k = std::exp(x[i]);  // quirk A.9-2: log10 limits, natural exp
      double xr = k * 8.0;
      double w = 3.0 * (std::sin(xr) - xr * std::cos(xr)) / (xr * xr * xr);
      rk[i] = k; rlnk[i] = x[i];
      rf[i] = rw[i] * (hi - lo) * (k * (k * w) * (k * w)) / (2.0 * M_PI * M_PI);
      rwn[i] = rw[i] * (hi - lo);  // jc_sigmasqr_f64: any R
    }
  }
  // ---- halofit grids ------------------------------------------------------------------------------
  std::vector<double> hlnk = linspace(std::log(1e-4), std::log(1e4), JC_NHFK), hk(JC_NHFK);
  for (int i = 0; i < JC_NHFK; ++i) hk[i] = std::exp(hlnk[i]);
  std::vector<double> hwk = simpson_weights(256, (std::log(1e4) - std::log(1e-4)) / 256);
  std::vector<double> hlogr = linspace(std::log(1e-4), std::log(1e1), JC_NHFR), hr(JC_NHFR);
  for (int i = 0; i < JC_NHFR; ++i) hr[i] = std::exp(hlogr[i]);
  // ---- lensing-efficiency grid ----------------------------------------------------------------------
  size_t nl = (size_t)JC_NLENS * JC_NLENS_COLS;
  std::vector<double> lens_t(n_src ? nl : 1), lens_z(n_src ? nl : 1);
  std::vector<uint16_t> lens_ix(n_src ? nl : 1);
  if (n_src) {
    for (int n = 0; n < JC_NLENS_COLS; ++n) {
      double delta = lens_zmax - lz[n];
      double step = delta / 256;
      for (int m = 0; m < JC_NLENS; ++m) {
        volatile double p = (double)m * step;
        double zp = (m == 256) ? lens_zmax : p + lz[n];
        double ap = 1.0 / (1.0 + zp);
        int i0, i1; double t;
        bracket(ap, atab, &i0, &i1, &t);
        size_t o = (size_t)m * JC_NLENS_COLS + n;
        lens_z[o] = zp; lens_t[o] = t; lens_ix[o] = (uint16_t)(i0 | (i1 << 8));
      }
    }
  }
  // ---- ell ----------------------------------------------------------------------------------------------
  std::vector<double> ell(ell_host, ell_host + L), ellp5(L), lnellp5(L), ellfac(L), covnorm(L), ell108(L), ell14(L), ellm3(L);
  for (int l = 0; l < L; ++l) {
    double e = ell[l];
    const double ep5 = grid_a ? e : e + 0.5;  // grid plan: the table entry is the wavenumber itself
    ellp5[l] = ep5;
    lnellp5[l] = std::log(ep5);
    ell108[l] = std::pow(ep5, 1.08);
    ell14[l] = std::pow(ep5, 1.4);
    ellm3[l] = 1.0 / (ep5 * ep5 * ep5);
    ellfac[l] = grid_a ? 1.0 : std::sqrt((e - 1) * e * (e + 1) * (e + 2)) / ((e + 0.5) * (e + 0.5));  // probes.py:73
    double g;  // np.gradient(ell), unit spacing (angular_cl.py:139)
    if (L == 1) g = 0.0;
    else if (l == 0) g = ell[1] - ell[0];
    else if (l == L - 1) g = ell[L - 1] - ell[L - 2];
    else g = (ell[l + 1] - ell[l - 1]) / 2.0;
    covnorm[l] = (2 * e + 1) * g;
  }
  // ---- pairs, bias tables, noise ---------------------------------------------------------------------------
  std::vector<uint8_t> pi(P), pj(P);
  { int p = 0; for (int i = 0; i < T; ++i) for (int j = i; j < T; ++j) { pi[p] = i; pj[p] = j; ++p; } }
  std::vector<double> bias_node((size_t)JC_NA_PAD * d.TS, 0.0);  // node-major [n][TS]
  for (int t = 0; t < T; ++t) {
    const jc_tracer& tr = pb->tracers[t];
    bool needs_bias = tr.kind == JC_TRACER_NUMBER_COUNTS || tr.ia_enabled;
    for (int n = 0; n < JC_NA && needs_bias; ++n) {
      double b = tr.bias.params[0];  // constant / inverse_growth: b (bias.py:20-22,37-39)
      if (tr.bias.family == JC_BIAS_DES_Y1_IA)  // bias.py:55-57
        b = tr.bias.params[0] * std::pow((1.0 + lz[n]) / (1.0 + tr.bias.params[2]), tr.bias.params[1]);
      bias_node[(size_t)n * d.TS + t] = b;
    }
  }

  jc_plan* plan = new jc_plan();
  memset(plan, 0, sizeof(*plan));
  plan->device = device;
  plan->problem = *pb;
  for (int t = 0; t < T; ++t) {
    const jc_tracer& tr = pb->tracers[t];
    double ng = tr.nz.gals_per_arcmin2 * JC_STERADIAN_TO_ARCMIN2;  // redshift.py:49-51
    plan->noise[t] = tr.kind == JC_TRACER_WEAK_LENSING ? tr.sigma_e * tr.sigma_e / ng : 1.0 / ng;
  }

  size_t o_chi_pt_a = B.add(chi_pt_a), o_chi_pt_lna = B.add(chi_pt_lna), o_chi_h6 = B.add(chi_h6);
  size_t o_gr_pt_a = B.add(gr_pt_a), o_gr_pt_lna = B.add(gr_pt_lna), o_gr_h = B.add(gr_h);
  size_t o_la = B.add(la), o_llna = B.add(llna), o_lz = B.add(lz), o_lw = B.add(lw);
  size_t o_lct = B.add(lct), o_lgt = B.add(lgt), o_lcix = B.add(lcix), o_lgix = B.add(lgix);
  size_t o_rk = B.add(rk), o_rlnk = B.add(rlnk), o_rf = B.add(rf), o_rwn = B.add(rwn);
  size_t o_hk = B.add(hk), o_hlnk = B.add(hlnk), o_hwk = B.add(hwk), o_hr = B.add(hr), o_hlogr = B.add(hlogr);
  size_t o_lens_t = B.add(lens_t), o_lens_ix = B.add(lens_ix), o_lens_z = B.add(lens_z);
  size_t o_lens_nw = B.reserve((size_t)(n_src ? n_src : 1) * nl * sizeof(double));
  size_t o_nz_node = B.reserve((size_t)JC_NA_PAD * d.TS * sizeof(double));
  size_t o_bias_node = B.add(bias_node);
  size_t o_kind = B.add(tr_kind), o_inv = B.add(tr_inv), o_ia = B.add(tr_ia), o_src = B.add(tr_src);
  std::vector<int> fin_idx;
  for (int t = 0; t < T; ++t)
    if (tr_kind[t] != JC_TRACER_WEAK_LENSING || tr_delta_ix[t] >= 0 || tr_ia[t]) fin_idx.push_back(t);
  d.n_fin = (int)fin_idx.size();
  if (fin_idx.empty()) fin_idx.push_back(0);
  size_t o_fin = B.add(fin_idx);
  size_t o_dix = B.add(tr_delta_ix), o_dt = B.add(tr_delta_t);
  std::vector<size_t> o_kde_z(T, 0), o_kde_w(T, 0);
  for (int t = 0; t < T; ++t) {
    const jc_nz& nz = pb->tracers[t].nz;
    if (nz.family == JC_NZ_KDE) {
      o_kde_z[t] = B.add(nz.kde_z, (size_t)nz.kde_n * sizeof(double));
      o_kde_w[t] = B.add(nz.kde_w, (size_t)nz.kde_n * sizeof(double));
    }
  }
  size_t o_m1 = B.add(tr_m1), o_srct = B.add(src_tracer);
  size_t o_ell = B.add(ell), o_ellp5 = B.add(ellp5), o_lnellp5 = B.add(lnellp5);
  size_t o_ellfac = B.add(ellfac), o_covnorm = B.add(covnorm);
  size_t o_ell108 = B.add(ell108), o_ell14 = B.add(ell14), o_ellm3 = B.add(ellm3);
  size_t o_pi = B.add(pi), o_pj = B.add(pj);
  // contraction order (filled after the n(z) kernels below): pairs sorted by the first Limber stage their kernel
  // product can be non-zero at, tiles of 8 pairs with the stage range each tile has to visit
  const int Ppad8 = (P + 7) & ~7, mtiles = Ppad8 / 8;
  size_t o_cpi = B.reserve(Ppad8), o_cpj = B.reserve(Ppad8), o_cpo = B.reserve((size_t)Ppad8 * sizeof(uint16_t));
  size_t o_tlo = B.reserve(mtiles), o_thi = B.reserve(mtiles);
  // tables of the kernels' table-driven exp / log (jc_math.cuh): 2^(j/256); {c_j, -ln c_j}
  std::vector<double> math_tab(JCM_TAB_DOUBLES);
  jc_math_table(math_tab.data());
  size_t o_math = B.add(math_tab);
  size_t o_norm = B.reserve(JC_MAX_TRACERS * sizeof(double));

  unsigned char* base = nullptr;
  cudaError_t ce = cudaMalloc(&base, B.host.size());
  if (ce != cudaSuccess) { jc_set_cuda_error(ce, "cudaMalloc(plan)"); delete plan; return JC_ERR_CUDA; }
  ce = cudaMemcpy(base, B.host.data(), B.host.size(), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { jc_set_cuda_error(ce, "cudaMemcpy(plan)"); cudaFree(base); delete plan; return JC_ERR_CUDA; }
  plan->dev_blob = base; plan->dev_blob_bytes = B.host.size();
#define DP(T_, off) ((const T_*)(base + (off)))
  d.chi_pt_a = DP(double, o_chi_pt_a); d.chi_pt_lna = DP(double, o_chi_pt_lna); d.chi_h6 = DP(double, o_chi_h6);
  d.gr_pt_a = DP(double, o_gr_pt_a); d.gr_pt_lna = DP(double, o_gr_pt_lna); d.gr_h = DP(double, o_gr_h);
  d.limb_a = DP(double, o_la); d.limb_lna = DP(double, o_llna); d.limb_z = DP(double, o_lz); d.limb_w = DP(double, o_lw);
  d.limb_chi_t = DP(double, o_lct); d.limb_gr_t = DP(double, o_lgt);
  d.limb_chi_ix = DP(uint16_t, o_lcix); d.limb_gr_ix = DP(uint16_t, o_lgix);
  d.romb_k = DP(double, o_rk); d.romb_lnk = DP(double, o_rlnk); d.romb_f = DP(double, o_rf); d.romb_w = DP(double, o_rwn);
  d.hf_k = DP(double, o_hk); d.hf_lnk = DP(double, o_hlnk); d.hf_wk = DP(double, o_hwk);
  d.hf_r = DP(double, o_hr); d.hf_logr = DP(double, o_hlogr);
  d.lens_t = DP(double, o_lens_t); d.lens_ix = DP(uint16_t, o_lens_ix); d.lens_nw = DP(double, o_lens_nw);
  d.nz_node = DP(double, o_nz_node); d.bias_node = DP(double, o_bias_node);
  d.tr_kind = DP(int, o_kind); d.tr_inv_growth = DP(int, o_inv); d.tr_ia = DP(int, o_ia); d.tr_src = DP(int, o_src);
  d.tr_delta_ix = DP(int, o_dix); d.tr_delta_t = DP(double, o_dt); d.fin_idx = DP(int, o_fin);
  d.tr_m1 = DP(double, o_m1); d.src_tracer = DP(int, o_srct);
  d.ell = DP(double, o_ell); d.ellp5 = DP(double, o_ellp5); d.lnellp5 = DP(double, o_lnellp5);
  d.lnl_min = *std::min_element(lnellp5.begin(), lnellp5.end());
  d.lnl_max = *std::max_element(lnellp5.begin(), lnellp5.end());
  d.lnl_step = 0.0;  // > 0: ln(ell + 1/2) is (nearly) uniformly spaced -- lets the power kernel snap its table spacing
  if (L >= 8) {      // median spacing, accepted when every spacing is within 10 % of it (np.logspace in ell: ell + 1/2 is not exactly log-uniform)
    std::vector<double> dl(L - 1);
    for (int l = 0; l + 1 < L; ++l) dl[l] = lnellp5[l + 1] - lnellp5[l];
    std::vector<double> sorted(dl);
    std::sort(sorted.begin(), sorted.end());
    const double med = sorted[(L - 1) / 2];
    bool uniform = med > 0.0;
    for (int l = 0; l + 1 < L && uniform; ++l) uniform = std::fabs(dl[l] - med) < 0.1 * med;
    if (uniform) d.lnl_step = med;
  }
  d.ellfac = DP(double, o_ellfac); d.covnorm = DP(double, o_covnorm);
  d.ell108 = DP(double, o_ell108); d.ell14 = DP(double, o_ell14); d.ellm3 = DP(double, o_ellm3);
  d.pair_i = DP(uint8_t, o_pi); d.pair_j = DP(uint8_t, o_pj);
  d.cpair_i = DP(uint8_t, o_cpi); d.cpair_j = DP(uint8_t, o_cpj); d.cpair_out = DP(uint16_t, o_cpo);
  d.ctile_lo = DP(uint8_t, o_tlo); d.ctile_hi = DP(uint8_t, o_thi);
  d.math_tab = DP(double, o_math);
  plan->d = d;

  // ---- one-time n(z) kernels ---------------------------------------------------------------------------------
  NzDevAll all;
  memset(&all, 0, sizeof(all));
  for (int t = 0; t < T; ++t) {
    const jc_nz& nz = pb->tracers[t].nz;
    all.nz[t].family = nz.family; all.nz[t].n_shifts = nz.n_shifts; all.nz[t].zmax = nz.zmax;
    for (int i = 0; i < 4; ++i) all.nz[t].p[i] = nz.params[i];
    for (int i = 0; i < JC_MAX_SHIFTS; ++i) all.nz[t].shifts[i] = nz.shifts[i];
    if (nz.family == JC_NZ_KDE) {
      all.nz[t].kde_z = (const double*)(base + o_kde_z[t]);
      all.nz[t].kde_w = (const double*)(base + o_kde_w[t]);
      all.nz[t].kde_n = nz.kde_n;
      all.nz[t].kde_bw = nz.kde_bw;
    }
  }
  double* norm = (double*)(base + o_norm);
  jc_nz_norm_kernel<<<T, 288>>>(all, norm);
  jc_nz_node_kernel<<<dim3((JC_NA + 127) / 128, T), 128>>>(all, norm, d.limb_z, (double*)(base + o_nz_node), d.TS);
  if (n_src)
    jc_nz_lens_kernel<<<dim3(JC_NLENS_COLS / 128, JC_NLENS, n_src), 128>>>(
        all, (const int*)(base + o_srct), norm, (const double*)(base + o_lens_z), (double*)(base + o_lens_nw));
  ce = cudaDeviceSynchronize();
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (ce != cudaSuccess) { jc_set_cuda_error(ce, "plan n(z) kernels"); cudaFree(base); delete plan; return JC_ERR_CUDA; }
  // ---- contraction supports --------------------------------------------------------------------------------
  // A number-counts kernel is n_i(z_n) b_i H (probes.py:77-99): zero wherever the cosmology-independent n_i(z_n) is.
  // Per tracer the node range outside of which |n_i| <= eps * max|n_i| (eps = 0: exactly zero -> bitwise the full sum;
  // the default 1e-20 drops contributions below 1e-20 of the bin's peak, far under one ulp of the sum); lensing
  // tracers keep every node.  K4 visits, per tile of 8 pairs, only the 12-node stages inside the tile's range.
  {
    std::vector<double> nzh((size_t)JC_NA_PAD * d.TS);
    ce = cudaMemcpy(nzh.data(), base + o_nz_node, nzh.size() * sizeof(double), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) { jc_set_cuda_error(ce, "cudaMemcpy(nz_node)"); cudaFree(base); delete plan; return JC_ERR_CUDA; }
    const double eps = g_jc_contract_eps;
    const int n_stage = (JC_NA + 11) / 12;
    std::vector<int> t_lo(T, 0), t_hi(T, JC_NA - 1);
    for (int t = 0; t < T && !grid_a; ++t) {
      if (tr_kind[t] != JC_TRACER_NUMBER_COUNTS) continue;
      double mx = 0.0;
      for (int n = 0; n < JC_NA; ++n) mx = std::max(mx, std::fabs(nzh[(size_t)n * d.TS + t]));
      int lo = JC_NA, hi = -1;
      for (int n = 0; n < JC_NA; ++n) {
        const double v = std::fabs(nzh[(size_t)n * d.TS + t]);
        if (!(v <= eps * mx) || (eps <= 0.0 && v != 0.0) || v != v) { if (n < lo) lo = n; hi = n; }
      }
      t_lo[t] = lo; t_hi[t] = hi;  // lo > hi: the kernel vanishes everywhere
    }
    std::vector<int> p_lo(P), p_hi(P), order(P);
    for (int q = 0; q < P; ++q) {
      const int lo = std::max(t_lo[pi[q]], t_lo[pj[q]]), hi = std::min(t_hi[pi[q]], t_hi[pj[q]]);
      if (lo > hi) { p_lo[q] = n_stage; p_hi[q] = -1; } else { p_lo[q] = lo / 12; p_hi[q] = hi / 12; }
      order[q] = q;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (p_lo[a] != p_lo[b]) return p_lo[a] < p_lo[b];
      return p_hi[a] > p_hi[b];
    });
    std::vector<uint8_t> cpi(Ppad8), cpj(Ppad8), tlo(mtiles), thi(mtiles);
    std::vector<uint16_t> cpo(Ppad8);
    for (int q = 0; q < Ppad8; ++q) {
      const int src = order[q < P ? q : P - 1];  // pad rows repeat the last pair and are never stored
      cpi[q] = pi[src]; cpj[q] = pj[src]; cpo[q] = (uint16_t)src;
    }
    for (int m = 0; m < mtiles; ++m) {
      int lo = n_stage, hi = -1;
      for (int q = 8 * m; q < std::min(8 * m + 8, P); ++q) { lo = std::min(lo, p_lo[order[q]]); hi = std::max(hi, p_hi[order[q]]); }
      if (hi < lo) { lo = 1; hi = 0; }  // empty: no stage
      if (eps < 0.0) { lo = 0; hi = n_stage - 1; }  // contract_eps < 0: every stage (A/B knob)
      tlo[m] = (uint8_t)lo; thi[m] = (uint8_t)hi;
    }
    ce = cudaMemcpy(base + o_cpi, cpi.data(), cpi.size(), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(base + o_cpj, cpj.data(), cpj.size(), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(base + o_cpo, cpo.data(), cpo.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(base + o_tlo, tlo.data(), tlo.size(), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(base + o_thi, thi.data(), thi.size(), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { jc_set_cuda_error(ce, "cudaMemcpy(contraction order)"); cudaFree(base); delete plan; return JC_ERR_CUDA; }
  }
#undef DP
  *plan_out = plan;
  return JC_OK;
}

extern "C" int jc_plan_create(const jc_problem* pb, const double* ell_host, int32_t n_ell,
                              int32_t device, jc_plan** plan_out) {
  return create_plan(pb, ell_host, n_ell, device, nullptr, 0, plan_out);
}

// Grid plan for the stand-alone background / matter-power functions (background.py, power.py of the reference): K1 and
// K3 run unchanged on the caller's scale factors and wavenumbers; tracer kernels and the contraction are not used.
extern "C" int jc_grid_plan_create(int32_t transfer, int32_t nonlinear, int32_t growth, const double* k_host, int32_t n_k,
                                   const double* a_host, int32_t n_a, int32_t device, jc_plan** plan_out) {
  if (!k_host || !a_host || n_k < 1 || n_a < 1 || n_a > JC_NA - 1) return JC_ERR_INVALID;
  for (int i = 0; i < n_k; ++i) if (!(k_host[i] > 0.0)) return JC_ERR_INVALID;
  for (int i = 0; i < n_a; ++i) if (!(a_host[i] > 0.0)) return JC_ERR_INVALID;
  jc_problem pb;
  memset(&pb, 0, sizeof(pb));
  pb.abi_version = JC_ABI_VERSION;
  pb.n_tracers = 1;  // placeholder tracer: its tables are built but never read in grid mode
  pb.transfer = transfer; pb.nonlinear = nonlinear; pb.growth = growth;
  jc_tracer& tr = pb.tracers[0];
  tr.kind = JC_TRACER_NUMBER_COUNTS;
  tr.nz.family = JC_NZ_SMAIL; tr.nz.params[0] = 2.0; tr.nz.params[1] = 2.0; tr.nz.params[2] = 1.0;
  tr.nz.gals_per_arcmin2 = 1.0; tr.nz.zmax = 10.0; tr.probe_zmax = 10.0;
  tr.bias.family = JC_BIAS_CONSTANT; tr.bias.params[0] = 1.0;
  return create_plan(&pb, k_host, n_k, device, a_host, n_a, plan_out);
}

// n(z) as a stand-alone call: redshift_distribution.__call__ (redshift.py:27-31) = pz_fn(z) / simps(pz_fn, 0, zmax, 256),
// with the same device functions the plan tables are built from.  Host pointers, synchronous.
namespace {
__global__ void jc_nz_eval_kernel(NzDevAll all, const double* __restrict__ norm, const double* __restrict__ z,
                                  double* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = pz_fn(all.nz[0], z[i]) / norm[0];
}
}  // namespace

extern "C" int jc_nz_eval_f64(const jc_nz* nz, const double* z_host, int64_t n, double* out_host) {
  if (!nz || !z_host || !out_host || n < 1) return JC_ERR_INVALID;
  if (nz->family < JC_NZ_SMAIL || nz->family > JC_NZ_KDE || nz->family == JC_NZ_DELTA) return JC_ERR_UNSUPPORTED;
  if (nz->n_shifts < 0 || nz->n_shifts > JC_MAX_SHIFTS || !(nz->zmax > 0.0)) return JC_ERR_INVALID;
  if (nz->family == JC_NZ_KDE && (!nz->kde_z || !nz->kde_w || nz->kde_n < 1 || !(nz->kde_bw > 0.0))) return JC_ERR_INVALID;
  const size_t kde_bytes = nz->family == JC_NZ_KDE ? (size_t)nz->kde_n * sizeof(double) : 0;
  double* buf = nullptr;  // [z n][out n][norm 1][kde_z][kde_w]
  JC_CUDA_TRY(cudaMalloc(&buf, (2 * (size_t)n + 1) * sizeof(double) + 2 * kde_bytes));
  double *dz = buf, *dout = buf + n, *dnorm = buf + 2 * n, *dkz = dnorm + 1, *dkw = dkz + (kde_bytes / sizeof(double));
  NzDevAll all;
  memset(&all, 0, sizeof(all));
  NzDev& d = all.nz[0];
  d.family = nz->family; d.n_shifts = nz->n_shifts; d.zmax = nz->zmax;
  for (int i = 0; i < 4; ++i) d.p[i] = nz->params[i];
  for (int i = 0; i < JC_MAX_SHIFTS; ++i) d.shifts[i] = nz->shifts[i];
  cudaError_t e = cudaMemcpy(dz, z_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && kde_bytes) {
    e = cudaMemcpy(dkz, nz->kde_z, kde_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dkw, nz->kde_w, kde_bytes, cudaMemcpyHostToDevice);
    d.kde_z = dkz; d.kde_w = dkw; d.kde_n = nz->kde_n; d.kde_bw = nz->kde_bw;
  }
  if (e == cudaSuccess) {
    jc_nz_norm_kernel<<<1, 288>>>(all, dnorm);
    jc_nz_eval_kernel<<<(unsigned)((n + 255) / 256), 256>>>(all, dnorm, dz, dout, (long long)n);
    e = cudaMemcpy(out_host, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
  }
  cudaFree(buf);
  if (e != cudaSuccess) { jc_set_cuda_error(e, "jc_nz_eval_f64"); return JC_ERR_CUDA; }
  return JC_OK;
}

// Grid plan over a real tracer set: the radial kernels of the probes (probes.py: WeakLensing.kernel / NumberCounts.kernel)
// at the caller's scale factors.  No wavenumbers: the power kernel is not used with this plan.
extern "C" int jc_grid_plan_create_probes(const jc_problem* pb, const double* a_host, int32_t n_a, int32_t device,
                                          jc_plan** plan_out) {
  if (!pb || !a_host || n_a < 1 || n_a > JC_NA - 1) return JC_ERR_INVALID;
  for (int i = 0; i < n_a; ++i) if (!(a_host[i] > 0.0)) return JC_ERR_INVALID;
  const double k_dummy = 1.0;
  return create_plan(pb, &k_dummy, 1, device, a_host, n_a, plan_out);
}

extern "C" void jc_plan_destroy(jc_plan* plan) {
  if (!plan) return;
  JcDeviceGuard guard(plan->device);
  if (plan->prof) {
    for (int i = 0; i < JC_PROF_SLOTS; ++i)
      for (int j = 0; j <= JC_N_STAGES; ++j) if (plan->prof->ev[i][j]) cudaEventDestroy(plan->prof->ev[i][j]);
    delete plan->prof;
  }
  if (plan->dev_blob) cudaFree(plan->dev_blob);
  if (plan->arena_ws) cudaFree(plan->arena_ws);
  if (plan->arena_cosmo) cudaFree(plan->arena_cosmo);
  for (int i = 0; i < 2; ++i) {
    if (plan->arena_cl[i]) cudaFree(plan->arena_cl[i]);
    if (plan->ev_done[i]) cudaEventDestroy(plan->ev_done[i]);
    if (plan->ev_copied[i]) cudaEventDestroy(plan->ev_copied[i]);
  }
  if (plan->s_compute) cudaStreamDestroy(plan->s_compute);
  if (plan->s_copy) cudaStreamDestroy(plan->s_copy);
  delete plan;
}

extern "C" int32_t jc_plan_n_tracers(const jc_plan* plan) { return plan ? plan->d.T : 0; }
extern "C" int32_t jc_plan_n_cls(const jc_plan* plan) { return plan ? plan->d.P : 0; }
extern "C" int32_t jc_plan_n_ell(const jc_plan* plan) { return plan ? plan->d.L : 0; }
extern "C" int32_t jc_plan_n_cosmo_params(const jc_plan* plan) { return plan ? plan->d.ncp : 0; }

extern "C" int jc_noise_f64(const jc_plan* plan, double* noise_host) {
  if (!plan || !noise_host) return JC_ERR_INVALID;
  for (int t = 0; t < plan->d.T; ++t) noise_host[t] = plan->noise[t];
  return JC_OK;
}
