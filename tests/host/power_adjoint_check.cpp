// Host check of csrc/jc_power_point.cuh: the reverse-mode gradient of the K3 point function against a complex-step derivative
// of its value (std::complex instantiation of jc_point_value), input by input, on random points of the physical ranges.
// Built and run by tests/test_power_adjoint_host.py (g++, no CUDA).  Prints "max_rel_err <x> max_val_diff <y> points <n>".
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>

#include "../../jax_cosmo_b200/csrc/jc_power_point.cuh"

typedef std::complex<double> cplx;

struct HostMath {
  template <class S> S log(S x) const { return std::log(x); }
  template <class S> S exp(S x) const { return std::exp(x); }
  template <class S> S expb(S x) const { return std::exp(x); }
  template <class S> S sin(S x) const { return std::sin(x); }
  template <class S> S rcp(S x) const { return 1.0 / x; }
  double rcbrt(double x) const { return 1.0 / std::cbrt(x); }
  cplx rcbrt(cplx x) const { return std::pow(x, -1.0 / 3.0); }
};

struct Grad {
  double node[32], scal[32], ell;
  Grad() { for (int i = 0; i < 32; ++i) node[i] = scal[i] = 0.0; ell = 0.0; }
};
struct GradAcc {
  Grad* G;
  void node(int f, double g) { G->node[f] += g; }
  void scal(int f, double g) { G->scal[f] += g; }
  void ell(double g) { G->ell += g; }
};

template <class S> S* field_ptr(JcPointIn<S>& in, int kind, int f) {
  if (kind == 0) switch (f) {
    case JCP_NODE_INVCHIC: return &in.invchic; case JCP_NODE_LNCHIC: return &in.lnchic; case JCP_NODE_RNL: return &in.rnl;
    case JCP_NODE_LNKNL: return &in.lnknl; case JCP_NODE_AN: return &in.an; case JCP_NODE_BN: return &in.bn;
    case JCP_NODE_LNCF: return &in.lncf; case JCP_NODE_P3: return &in.p3; case JCP_NODE_ALPHA: return &in.alpha;
    case JCP_NODE_BETA: return &in.beta; case JCP_NODE_NU: return &in.nu; case JCP_NODE_E1: return &in.e1;
    case JCP_NODE_E2: return &in.e2; case JCP_NODE_NQ108: return &in.nq108; case JCP_NODE_NSILK: return &in.nsilk;
    case JCP_NODE_NAMP: return &in.namp; case JCP_NODE_GK: return &in.gk; case JCP_NODE_MU: return &in.mu;
    default: return nullptr;
  }
  if (kind == 1) switch (f) {
    case JCP_SCAL_INV13KEQ: return &in.inv13keq; case JCP_SCAL_BETA_C: return &in.beta_c; case JCP_SCAL_C14_ALPHA_C: return &in.c14ac;
    case JCP_SCAL_SH_D: return &in.shd; case JCP_SCAL_ALPHA_B: return &in.alpha_b; case JCP_SCAL_BETA_B: return &in.beta_b;
    case JCP_SCAL_BETA_NODE: return &in.beta_node; case JCP_SCAL_FB: return &in.fb; case JCP_SCAL_FC: return &in.fc;
    default: return nullptr;
  }
  return f == 0 ? &in.lpns : nullptr;
}

template <class S> JcPointIn<S> make_point(std::mt19937_64& rng) {
  std::uniform_real_distribution<double> U(0.0, 1.0);
  auto lu = [&](double lo, double hi) { return std::exp(std::log(lo) + U(rng) * (std::log(hi) - std::log(lo))); };
  auto un = [&](double lo, double hi) { return lo + U(rng) * (hi - lo); };
  JcPointIn<S> in;
  const double lp5 = lu(10.5, 1.0e4), chi = lu(1.0, 6.0e3), ns = un(0.9, 1.0);
  const double keq = un(0.008, 0.02), ksilk = un(0.06, 0.2);
  in.lp5 = lp5; in.lnl = std::log(lp5); in.l108 = std::pow(lp5, 1.08); in.l14 = std::pow(lp5, 1.4); in.lm3 = 1.0 / (lp5 * lp5 * lp5);
  in.lpns = std::pow(lp5, 3.0 + ns);
  in.invchic = 1.0 / chi; in.lnchic = std::log(chi);
  in.nq108 = std::pow(13.41 * keq * chi, -1.08); in.nsilk = std::pow(ksilk * chi, -1.4);
  in.namp = lu(1e-2, 1e3) / std::pow(lp5, 3.0 + ns);  // lpns namp = Delta^2_L / T^2 in [1e-2, 1e3]
  in.rnl = lu(0.05, 20.0); in.lnknl = -std::log(in.rnl);
  in.beta = un(0.5, 3.0); in.alpha = un(1.0, 8.0); in.e1 = un(2.5, 3.2); in.e2 = un(0.8, 1.1); in.p3 = un(2.0, 2.9);
  in.lncf = un(-2.0, 3.0); in.an = lu(0.1, 10.0); in.nu = lu(0.01, 100.0); in.mu = lu(0.001, 0.1); in.bn = lu(0.1, 5.0);
  in.gk = lu(1.0, 1e4);
  in.inv13keq = 1.0 / (13.41 * keq); in.beta_c = un(0.5, 0.9); in.c14ac = 14.2 / un(0.5, 1.0); in.shd = un(90.0, 160.0);
  in.alpha_b = un(0.3, 1.0); in.beta_b = un(0.5, 1.2); in.beta_node = un(3.0, 7.0);
  in.fb = un(0.1, 0.25); in.fc = 1.0 - un(0.1, 0.25);
  return in;
}

template <class A, class B> void copy_in(const JcPointIn<A>& a, JcPointIn<B>& b) {
  b.lp5 = a.lp5; b.lnl = a.lnl; b.l108 = a.l108; b.l14 = a.l14; b.lm3 = a.lm3; b.lpns = a.lpns;
  b.invchic = a.invchic; b.lnchic = a.lnchic; b.nq108 = a.nq108; b.nsilk = a.nsilk; b.namp = a.namp; b.rnl = a.rnl; b.lnknl = a.lnknl;
  b.beta = a.beta; b.alpha = a.alpha; b.e1 = a.e1; b.e2 = a.e2; b.p3 = a.p3; b.lncf = a.lncf; b.an = a.an; b.nu = a.nu; b.mu = a.mu;
  b.bn = a.bn; b.gk = a.gk; b.inv13keq = a.inv13keq; b.beta_c = a.beta_c; b.c14ac = a.c14ac; b.shd = a.shd; b.alpha_b = a.alpha_b;
  b.beta_b = a.beta_b; b.beta_node = a.beta_node; b.fb = a.fb; b.fc = a.fc;
}

int main(int argc, char** argv) {
  const int n_points = argc > 1 ? atoi(argv[1]) : 2000;
  std::mt19937_64 rng(20240607);
  HostMath m;
  double worst = 0.0, worst_val = 0.0;
  int worst_kind = -1, worst_f = -1, worst_mode = -1;
  for (int p = 0; p < n_points; ++p) {
    JcPointIn<double> in = make_point<double>(rng);
    for (int mode = 0; mode < 3; ++mode) {  // linear, takahashi2012, smith2003
      const bool halofit = mode > 0, smith = mode == 2;
      Grad G;
      GradAcc acc{&G};
      const double V = jc_point_adjoint(in, m, halofit, smith, acc);
      const double Vv = jc_point_value<double>(in, m, halofit, smith);
      worst_val = std::fmax(worst_val, std::fabs(V - Vv) / std::fabs(Vv));
      // scale of the gradient terms: |x_j dV/dx_j| summed (a relative error per input would blow up at sign changes)
      for (int kind = 0; kind < 3; ++kind)
        for (int f = 0; f < 32; ++f) {
          JcPointIn<cplx> ic;
          copy_in(in, ic);
          cplx* x = field_ptr(ic, kind, f);
          if (!x) continue;
          const double x0 = x->real(), h = 1e-30 * std::fmax(std::fabs(x0), 1e-3);
          *x = cplx(x0, h);
          const double dcs = jc_point_value<cplx>(ic, m, halofit, smith).imag() / h;
          const double dad = kind == 0 ? G.node[f] : (kind == 1 ? G.scal[f] : G.ell);
          const double scale = std::fmax(std::fabs(dcs), std::fabs(V / x0) * 1e-3);
          const double err = std::fabs(dad - dcs) / scale;
          if (err > worst) { worst = err; worst_kind = kind; worst_f = f; worst_mode = mode; }
        }
    }
  }
  std::printf("max_rel_err %.3e max_val_diff %.3e points %d worst_input kind=%d field=%d mode=%d\n", worst, worst_val, n_points,
              worst_kind, worst_f, worst_mode);
  return 0;
}
