"""The derivative oracle pins itself on the CPU: the complex-step Jacobian of the restatement (oracle/derivatives.py::
cs_jacobian -- decisions on real parts, i.e. the frozen-index derivative jax.jacfwd returns for the reference) against
index-checked 4th-order finite differences, and against two exact identities."""
import numpy as np

from oracle import cl_oracle as o
from oracle import derivatives as od
from oracle import scenarios as sc


def test_complex_step_vs_finite_differences():
    scn = sc.scenario("cs", sc.WCDM, sc.ELL_CFG2[::16], [sc.sources(5, 2.0, True), sc.lenses(5, 2.0, True)])
    prob, ell, row = sc.flatten_spec(scn), np.array(scn["ell"]), sc.cosmo_row(sc.WCDM)
    cl, jac = od.cs_jacobian(row, ell, prob)
    assert np.array_equal(cl, o.angular_cl(row, ell, prob))  # the real path is untouched by the complex-step helpers
    _, jfd, _ = od.fd_jacobian(row, ell, prob)
    scale = np.abs(jfd).max(axis=2, keepdims=True)
    worst = (np.abs(jac - jfd) / scale).reshape(7, -1).max(axis=1)
    assert worst[:5].max() < 1e-9, worst   # Omega_c, Omega_b, h, n_s, sigma8: FD noise ~1e-10
    assert worst.max() < 1e-6, worst       # w0, wa: FD noise 1e-9 ... 1e-7 (h relative to |theta|, theta = 0 for wa)


def test_complex_step_identities():
    # linear P(k): C_ell is exactly proportional to sigma8^2 (power.py:47) -> dC/dsigma8 = 2 C / sigma8 to rounding
    scn = sc.scenario("lin", sc.PLANCK15, sc.ELL_CFG1[::10], [sc.sources(4, 6.5)], "linear")
    prob, ell, row = sc.flatten_spec(scn), np.array(scn["ell"]), sc.cosmo_row(sc.PLANCK15)
    cl, jac = od.cs_jacobian(row, ell, prob, params=("sigma8", "n_s"))
    assert np.max(np.abs(jac[0] / (2.0 * cl / row[4]) - 1.0)) < 1e-12
    assert np.all(np.isfinite(jac)) and np.abs(jac[1]).max() > 0


def test_complex_step_gamma_growth():
    """9-column rows (growth index gamma, core.py:104-105): a perturbed gamma reaches the tracer kernels through the growth factor
    only (IA and inverse-growth bias), while H(z) and chi stay real -- the kernel table must still hold the imaginary part."""
    scn = [s for s in sc.golden_scenarios() if s["name"] == "switch_gamma_growth"][0]
    row, prob, ell = sc.cosmo_row(scn["cosmo"]), sc.flatten_spec(scn), sc.ELL_CFG2[::12]
    assert row.shape == (9,)
    cl, jac = od.cs_jacobian(row, ell, prob, params=("sigma8", "gamma"))
    assert np.array_equal(cl, o.angular_cl(row, ell, prob))
    _, jfd, _ = od.fd_jacobian(row, ell, prob, params=("gamma",))
    scale = np.abs(jac[1]).max(axis=1, keepdims=True)
    assert np.abs(jac[1]).max() > 0
    assert (np.abs(jfd[0] - jac[1]) / scale).max() < 1e-6
