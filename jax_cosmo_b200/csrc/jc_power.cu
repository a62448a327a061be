// jc_power.cu -- K3: V[n,l] = geom_n * P(k = (l+1/2)/max(chi_n,1), a_n) for every (node, ell).
#include "jc_internal.cuh"

namespace {

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

}  // namespace

void jc_launch_power(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  jc_power_kernel<<<dim3((JC_NA * pl.L + 255) / 256, chunk), 256, 0, s>>>(pl, ws);
}
