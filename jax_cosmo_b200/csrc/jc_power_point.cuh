// jc_power_point.cuh -- the (ell, node) point function of K3 with its reverse-mode gradient.
//
//   V = geom * P(k, a):   T(k) Eisenstein & Hu with wiggles (transfer.py:113-153), Delta^2_L (power.py:49-52),
//                         halofit takahashi2012 / smith2003 (power.py:246-262) -- the arithmetic of jc_power_kernel.
//
// Forward-mode passes (jc_dual.cuh) pay ~225 FP64 instructions per point and DIRECTION on top of the 186 of the value
// (ncu, profiles/r02_ncu_summary.md section 4: 71 % of the FP64 pipe, so it is the arithmetic).  V is a scalar function of
// 28 inputs whose tangents are per-node, per-ell or per-cosmology tables; one reverse sweep (~2 x the value) yields
// dV/d(input) for all of them, and every direction then costs one multiply-add per input with its table entry:
//     dV_k = sum_j  (dV / d in_j) * d(in_j)_k.
// `Acc` receives each gradient as soon as the sweep has it (no 28-entry gradient vector is ever live):
//     acc.node(field, g)   per-node workspace field JC_NODE_*        acc.ell(g)   ws.ellpow = (l+1/2)^(3+n_s)
//     acc.scal(field, g)   per-cosmology constant JC_SCAL_*
// `M` supplies log / exp / sin / rcbrt / rcp (device: jc_math.cuh; host: libm, tests/test_power_adjoint_host.py drives the
// sweep on the CPU against a complex-step derivative of jc_point_value below).
#pragma once

#if defined(__CUDACC__)
#define JCP_HD __device__ __forceinline__
#else
#define JCP_HD inline
#endif

// node / scal field numbers as in include/jc_b200.h (the host test includes this header without the CUDA headers)
#ifndef JC_B200_H_FIELDS
enum { JCP_NODE_INVCHIC = 1, JCP_NODE_LNCHIC = 2, JCP_NODE_RNL = 7, JCP_NODE_LNKNL = 8, JCP_NODE_AN = 11, JCP_NODE_BN = 12,
       JCP_NODE_LNCF = 13, JCP_NODE_P3 = 14, JCP_NODE_ALPHA = 15, JCP_NODE_BETA = 16, JCP_NODE_NU = 17, JCP_NODE_E1 = 18,
       JCP_NODE_E2 = 19, JCP_NODE_NQ108 = 20, JCP_NODE_NSILK = 21, JCP_NODE_NAMP = 22, JCP_NODE_GK = 23, JCP_NODE_MU = 24 };
enum { JCP_SCAL_INV13KEQ = 1, JCP_SCAL_BETA_C = 2, JCP_SCAL_C14_ALPHA_C = 3, JCP_SCAL_SH_D = 4, JCP_SCAL_ALPHA_B = 6,
       JCP_SCAL_BETA_B = 7, JCP_SCAL_BETA_NODE = 8, JCP_SCAL_FB = 9, JCP_SCAL_FC = 10 };
#endif

template <class S>
struct JcPointIn {
  // ell side (no derivative): l + 1/2, ln(l + 1/2), (l+1/2)^1.08, (l+1/2)^1.4, (l+1/2)^-3
  double lp5, lnl, l108, l14, lm3;
  S lpns;  // (l+1/2)^(3+n_s), ws.ellpow
  // node side (ws.node fields)
  S invchic, lnchic, nq108, nsilk, namp, rnl, lnknl, beta, alpha, e1, e2, p3, lncf, an, nu, mu, bn, gk;
  // cosmology side (ws.scal fields)
  S inv13keq, beta_c, c14ac, shd, alpha_b, beta_b, beta_node, fb, fc;
};

struct JcPointK {  // literal constants of the fits
  static constexpr double e1c = 2.718281828459045, c699 = 69.9, c142 = 14.2, c386 = 386.0, c18 = 1.8, inv54 = 1.0 / 5.4,
                          inv52 = 1.0 / 5.2, eighth = 0.125, quarter = 0.25, half_pi = 1.5707963267948966;
};

// Value only, any scalar type (double, std::complex<double> for the host's complex-step check).  Same operation order as
// eh_point<T, false> and the halofit block of jc_power_kernel.
template <class S, class M>
JCP_HD S jc_point_value(const JcPointIn<S>& in, const M& m, bool halofit, bool smith) {
  typedef JcPointK C;
  const S lnk = in.lnl - in.lnchic;
  const S k = in.lp5 * in.invchic;
  const S q108 = in.l108 * in.nq108, ks14 = in.l14 * in.nsilk;
  const S b18 = C::c18 * in.beta_c;
  const S bnode3 = in.beta_node * in.beta_node * in.beta_node, bb3 = in.beta_b * in.beta_b * in.beta_b;
  const S q = k * in.inv13keq;
  const S q2 = q * q;
  const S W = C::c699 * q108 + 1.0;
  const S U1 = C::c142 * W + C::c386;
  const S U2 = in.c14ac * W + C::c386;
  const S L1 = m.log(b18 * q + C::e1c);
  const S L2 = m.log(C::c18 * q + C::e1c);
  const S L1W = L1 * W, L2W = L2 * W;
  const S N1 = U1 * q2 + L1W, N2 = U2 * q2 + L1W, N3 = U1 * q2 + L2W;
  const S ks = k * in.shd;
  const S x54 = ks * C::inv54;
  const S x54_2 = x54 * x54;
  const S Fm1 = x54_2 * x54_2;
  const S numC = L1W * (Fm1 * N1 + N2);
  const S denC = (1.0 + Fm1) * (N1 * N2);
  const S ks2 = ks * ks, ks3 = ks2 * ks;
  const S arg = ks2 * m.rcbrt(ks3 + bnode3);
  const S x52 = ks * C::inv52;
  const S X52 = x52 * x52 + 1.0;
  const S BB = ks3 + bb3;
  const S silk = m.exp(-ks14);
  const S N3X = N3 * X52;
  const S numB = (L2W * BB + in.alpha_b * ks3 * silk * N3X) * m.sin(arg);
  const S denB = N3X * BB * arg;
  const S T = ((in.fb * numB) * denC + in.fc * numC * denB) * m.rcp(denB * denC);
  const S d2l = in.lpns * in.namp * (T * T);
  S d2 = d2l;
  if (halofit) {
    const S y = k * in.rnl;
    const S lny = lnk - in.lnknl;
    const S y2 = y * y;
    const S Nq = d2l * m.exp(in.beta * m.log(1.0 + d2l) - (y2 * C::eighth + C::quarter * y));
    const S Dq = in.alpha * d2l + 1.0;
    const S ye1 = m.expb(in.e1 * lny);
    const S ye2 = m.expb(in.e2 * lny);
    const S cfy = m.expb(in.p3 * (in.lncf + lny));
    const S Nh = in.an * ye1 * y2;
    S ynu = y2 + in.nu;
    if (smith) ynu = ynu + in.mu * y;
    const S Dh = (in.bn * ye2 + 1.0 + cfy) * ynu;
    d2 = (Nq * Dh + Nh * Dq) * m.rcp(Dq * Dh);
  }
  return d2 * in.lm3 * in.gk;
}

// Value and gradient: forward sweep as above (same operation order, so the value is the exact kernel's), then the reverse
// sweep.  Returns V; the gradient goes to `acc` input by input.
template <class M, class Acc>
JCP_HD double jc_point_adjoint(const JcPointIn<double>& in, const M& m, bool halofit, bool smith, Acc& acc) {
  typedef JcPointK C;
  // ---- forward ---------------------------------------------------------------------------------------------------------
  const double lnk = in.lnl - in.lnchic;
  const double k = in.lp5 * in.invchic;
  const double q108 = in.l108 * in.nq108, ks14 = in.l14 * in.nsilk;
  const double b18 = C::c18 * in.beta_c;
  const double bn2 = in.beta_node * in.beta_node, bnode3 = bn2 * in.beta_node;
  const double bb2 = in.beta_b * in.beta_b, bb3 = bb2 * in.beta_b;
  const double q = k * in.inv13keq;
  const double q2 = q * q;
  const double W = C::c699 * q108 + 1.0;
  const double U1 = C::c142 * W + C::c386;
  const double U2 = in.c14ac * W + C::c386;
  const double A1 = b18 * q + C::e1c, A2 = C::c18 * q + C::e1c;
  const double L1 = m.log(A1), L2 = m.log(A2);
  const double L1W = L1 * W, L2W = L2 * W;
  const double U1q2 = U1 * q2;
  const double N1 = U1q2 + L1W, N2 = U2 * q2 + L1W, N3 = U1q2 + L2W;
  const double ks = k * in.shd;
  const double x54 = ks * C::inv54;
  const double x54_2 = x54 * x54;
  const double Fm1 = x54_2 * x54_2;
  const double Mc = Fm1 * N1 + N2;
  const double numC = L1W * Mc;
  const double N1N2 = N1 * N2;
  const double denC = (1.0 + Fm1) * N1N2;
  const double ks2 = ks * ks, ks3 = ks2 * ks;
  const double cb = m.rcbrt(ks3 + bnode3);
  const double arg = ks2 * cb;
  const double x52 = ks * C::inv52;
  const double X52 = x52 * x52 + 1.0;
  const double BB = ks3 + bb3;
  const double silk = m.exp(-ks14);
  const double N3X = N3 * X52;
  const double sn = m.sin(arg), cs = m.sin(arg + C::half_pi);
  const double P2 = in.alpha_b * ks3;
  const double P3 = P2 * silk;
  const double numB0 = L2W * BB + P3 * N3X;
  const double numB = numB0 * sn;
  const double N3XBB = N3X * BB;
  const double denB = N3XBB * arg;
  const double fbnumB = in.fb * numB, fcnumC = in.fc * numC;
  const double rT = m.rcp(denB * denC);
  const double T = (fbnumB * denC + fcnumC * denB) * rT;
  const double TT = T * T;
  const double lpnamp = in.lpns * in.namp;
  const double d2l = lpnamp * TT;
  double d2 = d2l;
  // halofit intermediates
  double y = 0, lny = 0, y2 = 0, X1 = 1, lg = 0, Eq = 0, Nq = 0, Dq = 1, ye1 = 0, ye2 = 0, cfy = 0, s3 = 0, Nh = 0, ynu = 1, G = 1,
         Dh = 1, rr = 1;
  if (halofit) {
    y = k * in.rnl;
    lny = lnk - in.lnknl;
    y2 = y * y;
    X1 = 1.0 + d2l;
    lg = m.log(X1);
    Eq = m.exp(in.beta * lg - (y2 * C::eighth + C::quarter * y));
    Nq = d2l * Eq;
    Dq = in.alpha * d2l + 1.0;
    ye1 = m.expb(in.e1 * lny);
    ye2 = m.expb(in.e2 * lny);
    s3 = in.lncf + lny;
    cfy = m.expb(in.p3 * s3);
    Nh = in.an * ye1 * y2;
    ynu = y2 + in.nu;
    if (smith) ynu = ynu + in.mu * y;
    G = in.bn * ye2 + 1.0 + cfy;
    Dh = G * ynu;
    rr = m.rcp(Dq * Dh);
    d2 = (Nq * Dh + Nh * Dq) * rr;
  }
  const double d2lm3 = d2 * in.lm3;
  const double V = d2lm3 * in.gk;

  // ---- reverse: x_b = dV / dx ----------------------------------------------------------------------------------------------
  acc.node(JCP_NODE_GK, d2lm3);
  const double d2_b = in.lm3 * in.gk;
  double d2l_b, k_b = 0.0;
  if (halofit) {
    const double num_b = d2_b * rr;
    const double den_b = -(d2_b * d2) * rr;
    const double Nq_b = num_b * Dh, Nh_b = num_b * Dq;
    const double Dh_b = num_b * Nq + den_b * Dq;
    const double Dq_b = num_b * Nh + den_b * Dh;
    const double G_b = Dh_b * ynu, ynu_b = Dh_b * G;
    acc.node(JCP_NODE_BN, G_b * ye2);
    const double ye2_b = G_b * in.bn, cfy_b = G_b;
    acc.node(JCP_NODE_NU, ynu_b);
    double y2_b = ynu_b, y_b = 0.0;
    if (smith) {
      acc.node(JCP_NODE_MU, ynu_b * y);
      y_b = ynu_b * in.mu;
    }
    const double Nh_y2 = Nh_b * y2;
    acc.node(JCP_NODE_AN, Nh_y2 * ye1);
    const double ye1_b = Nh_y2 * in.an;
    y2_b = y2_b + Nh_b * (in.an * ye1);
    const double c3 = cfy_b * cfy;  // d/d(p3 s3)
    acc.node(JCP_NODE_P3, c3 * s3);
    const double s3_b = c3 * in.p3;
    acc.node(JCP_NODE_LNCF, s3_b);
    const double c2 = ye2_b * ye2, c1 = ye1_b * ye1;
    acc.node(JCP_NODE_E2, c2 * lny);
    acc.node(JCP_NODE_E1, c1 * lny);
    const double lny_b = s3_b + c2 * in.e2 + c1 * in.e1;
    acc.node(JCP_NODE_ALPHA, Dq_b * d2l);
    const double ex_b = (Nq_b * d2l) * Eq;
    acc.node(JCP_NODE_BETA, ex_b * lg);
    d2l_b = Dq_b * in.alpha + Nq_b * Eq + (ex_b * in.beta) * m.rcp(X1);
    y2_b = y2_b - ex_b * C::eighth;
    y_b = y_b - ex_b * C::quarter + 2.0 * y * y2_b;
    k_b = y_b * in.rnl;
    acc.node(JCP_NODE_RNL, y_b * k);
    acc.node(JCP_NODE_LNKNL, -lny_b);
    acc.node(JCP_NODE_LNCHIC, -lny_b);  // lnk = lnl - lnchic reaches V through ln y only
  } else {
    d2l_b = d2_b;
  }
  const double d2l_TT = d2l_b * TT;
  acc.ell(d2l_TT * in.namp);
  acc.node(JCP_NODE_NAMP, d2l_TT * in.lpns);
  const double T_b = d2l_b * lpnamp * (2.0 * T);
  // T = (fbnumB denC + fcnumC denB) rT,  rT = 1 / (denB denC)
  const double Tn_b = T_b * rT;
  const double Td_b = -(T_b * T) * rT;
  double denB_b = Td_b * denC + Tn_b * fcnumC;
  double denC_b = Td_b * denB + Tn_b * fbnumB;
  const double TnC = Tn_b * denC, TnB = Tn_b * denB;
  acc.scal(JCP_SCAL_FB, TnC * numB);
  acc.scal(JCP_SCAL_FC, TnB * numC);
  const double numB_b = TnC * in.fb, numC_b = TnB * in.fc;
  // denB = (N3X BB) arg ; numB = numB0 sn
  double N3X_b = denB_b * (BB * arg);
  double BB_b = denB_b * (N3X * arg);
  double arg_b = denB_b * N3XBB + (numB_b * numB0) * cs;
  const double numB0_b = numB_b * sn;
  // numB0 = L2W BB + P3 N3X ; P3 = (alpha_b ks3) silk
  double L2W_b = numB0_b * BB;
  BB_b = BB_b + numB0_b * L2W;
  N3X_b = N3X_b + numB0_b * P3;
  const double P3_b = numB0_b * N3X;
  const double P2_b = P3_b * silk;
  acc.node(JCP_NODE_NSILK, -(P3_b * P2) * silk * in.l14);  // silk = exp(-ks14), ks14 = l14 nsilk
  acc.scal(JCP_SCAL_ALPHA_B, P2_b * ks3);
  double ks3_b = P2_b * in.alpha_b + BB_b;
  acc.scal(JCP_SCAL_BETA_B, BB_b * (3.0 * bb2));
  // N3X = N3 X52 ; X52 = x52^2 + 1
  const double N3_b = N3X_b * X52;
  double ks_b = (N3X_b * N3) * (2.0 * x52 * C::inv52);
  // arg = ks2 cb ; cb = (ks3 + bnode3)^(-1/3)
  double ks2_b = arg_b * cb;
  const double cb2 = cb * cb;
  const double S3_b = (arg_b * ks2) * ((-1.0 / 3.0) * cb2 * cb2);
  ks3_b = ks3_b + S3_b;
  acc.scal(JCP_SCAL_BETA_NODE, S3_b * (3.0 * bn2));
  // ks3 = ks2 ks ; ks2 = ks ks
  ks2_b = ks2_b + ks3_b * ks;
  ks_b = ks_b + ks3_b * ks2 + ks2_b * (2.0 * ks);
  // denC = (1 + Fm1) N1N2 ; numC = L1W Mc ; Mc = Fm1 N1 + N2
  const double N1N2_b = denC_b * (1.0 + Fm1);
  const double Mc_b = numC_b * L1W;
  const double Fm1_b = denC_b * N1N2 + Mc_b * N1;
  const double N1_b = N1N2_b * N2 + Mc_b * Fm1;
  const double N2_b = N1N2_b * N1 + Mc_b;
  double L1W_b = numC_b * Mc + N1_b + N2_b;
  L2W_b = L2W_b + N3_b;
  ks_b = ks_b + Fm1_b * (4.0 * x54_2 * x54 * C::inv54);
  // N1 = U1 q2 + L1W ; N2 = U2 q2 + L1W ; N3 = U1 q2 + L2W
  const double U1_b = (N1_b + N3_b) * q2, U2_b = N2_b * q2;
  const double q2_b = (N1_b + N3_b) * U1 + N2_b * U2;
  // L1W = L1 W ; L2W = L2 W ; L1 = log A1 ; L2 = log A2
  const double A1_b = (L1W_b * W) * m.rcp(A1), A2_b = (L2W_b * W) * m.rcp(A2);
  double W_b = L1W_b * L1 + L2W_b * L2 + U1_b * C::c142 + U2_b * in.c14ac;
  acc.scal(JCP_SCAL_BETA_C, (A1_b * q) * C::c18);
  acc.scal(JCP_SCAL_C14_ALPHA_C, U2_b * W);
  acc.node(JCP_NODE_NQ108, (W_b * C::c699) * in.l108);  // W = 69.9 q108 + 1, q108 = l108 nq108
  const double q_b = A1_b * b18 + A2_b * C::c18 + q2_b * (2.0 * q);
  acc.scal(JCP_SCAL_INV13KEQ, q_b * k);
  acc.scal(JCP_SCAL_SH_D, ks_b * k);
  k_b = k_b + q_b * in.inv13keq + ks_b * in.shd;
  acc.node(JCP_NODE_INVCHIC, k_b * in.lp5);  // k = lp5 invchic
  return V;
}
