"""Derivative oracle for BASELINE config 4 (C_ell plus d/d theta of the 7 wCDM parameters).

TEST INFRASTRUCTURE ONLY.  The reference obtains these with `jax.jacfwd` (docs/notebooks/
jax-cosmo-intro.ipynb:989), i.e. the derivative of the *discretised* program with every nearest-node /
bracket index frozen.  Without JAX the oracle uses finite differences of the NumPy restatement
(oracle/cl_oracle.py) with the safeguard SURVEY 8(c) prescribes: C_ell(theta) jumps by ~2e-5 wherever
the halofit root index switches, so a stencil is only accepted when the index vector is identical at
every stencil point; otherwise the step is halved.  (All other interpolation brackets sit on fixed
grids and cannot switch.)

Stencil: 4th-order central difference (f(-2h) - 8 f(-h) + 8 f(h) - f(2h)) / 12h with
h = 3e-5 max(|theta|, 0.5).  The restatement's own evaluation noise is ~1e-13 relative, so the
round-off error is ~1e-13 / 3e-5 = 3e-9 of |C_ell| per unit theta and the truncation error
(h^4 f^(5) / 30) is negligible; a plain 2-point difference with h = 1e-6 |theta| is noise dominated
for parameters whose fiducial value is 0 (wa) -- measured 1e-5.
"""
import numpy as np

from oracle import cl_oracle as o


PARAM_INDEX = {"Omega_c": 0, "Omega_b": 1, "h": 2, "n_s": 3, "sigma8": 4, "Omega_k": 5, "w0": 6, "wa": 7, "gamma": 8}
WCDM_PARAMS = ("Omega_c", "Omega_b", "h", "n_s", "sigma8", "w0", "wa")



def cs_jacobian(cosmo_row, ell, problem, params=None, h=1e-30):
    """Complex-step derivative of the restatement: d cl / d theta_k = Im cl(theta + i h e_k) / h, exact to rounding (no
    subtraction, h = 1e-30), with every index / clip / abs decision of oracle/cl_oracle.py taken on real parts --
    the derivative of the discretised program with its indices frozen, which is what jax.jacfwd gives for the reference
    (docs/notebooks/jax-cosmo-intro.ipynb:989) and what the CUDA JVP kernels compute.  This is the primary derivative
    oracle (element-wise, ~1e-13); fd_jacobian below is the independent cross-check.
    -> (cl [P, L], jac [n_params, P, L])."""
    row = np.asarray(cosmo_row, dtype=np.float64)
    if params is None:
        params = WCDM_PARAMS + (("gamma",) if len(row) > 8 else ())
    cl0 = o.angular_cl(row, ell, problem)
    jac = np.empty((len(params),) + cl0.shape)
    for k, name in enumerate(params):
        r = row.astype(np.complex128)
        r[PARAM_INDEX[name]] += 1j * h
        out = o.angular_cl(r, ell, problem)
        if np.max(np.abs(out.real - cl0)) > 1e-13 * np.max(np.abs(cl0)):
            raise RuntimeError("complex-step evaluation changed the real part: a decision was not taken on real parts")
        jac[k] = out.imag / h
    return cl0, jac

def _eval(row, ell, problem):
    st = {}
    cl = o.angular_cl(row, ell, problem, stages=st)
    return cl, (st["root_ind"].copy() if "root_ind" in st else None)


def fd_jacobian(cosmo_row, ell, problem, params=WCDM_PARAMS, rel_step=3e-5, max_halvings=8):
    """-> (cl [P,L], jac [n_params,P,L], steps [n_params]).  Raises if no admissible step is found."""
    row = np.asarray(cosmo_row, dtype=np.float64)
    cl0, ind0 = _eval(row, ell, problem)
    jac = np.empty((len(params),) + cl0.shape)
    steps = np.empty(len(params))
    for k, name in enumerate(params):
        i = PARAM_INDEX[name]
        h = rel_step * max(abs(row[i]), 0.5)
        for _ in range(max_halvings):
            vals, ok = {}, True
            for m in (-2, -1, 1, 2):
                r = row.copy()
                r[i] = row[i] + m * h
                vals[m], ind = _eval(r, ell, problem)
                if ind0 is not None and not np.array_equal(ind, ind0):
                    ok = False
                    break
            if ok:
                jac[k] = (vals[-2] - 8.0 * vals[-1] + 8.0 * vals[1] - vals[2]) / (12.0 * h)
                steps[k] = h
                break
            h *= 0.5
        else:
            raise RuntimeError("no finite-difference step keeps the halofit root indices fixed for %s" % name)
    return cl0, jac, steps
