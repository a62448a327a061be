"""Generate tests/golden/*.npz by executing the UNMODIFIED reference source
(/root/reference/jax_cosmo) on the NumPy jax shim (oracle/jax_shim), float64.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (the GPU box has no
/root/reference); the .npz fixtures it writes are committed.  Usage:

    python oracle/make_golden.py [--only NAME] [--stages]
"""
import argparse
import json
import os
import sys
import time
import warnings

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "jax_shim"), "/root/reference", os.path.dirname(HERE)]

import numpy as np  # noqa: E402

import jax_cosmo as jc  # noqa: E402  (the reference, on the shim)
from jax_cosmo.angular_cl import (angular_cl, gaussian_cl_covariance,  # noqa: E402
                                  gaussian_cl_covariance_and_mean, noise_cl)

from oracle import scenarios as sc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def run_scenario(scn):
    cosmo = sc.build_cosmo(scn, jc)
    probes = sc.build_probes(scn, jc)
    tf, nl = sc.build_fns(scn, jc)
    ell = np.array(scn["ell"])
    t = time.time()
    cl = np.asarray(angular_cl(cosmo, ell, probes, transfer_fn=tf, nonlinear_fn=nl))
    noise = np.asarray(noise_cl(ell, probes))
    cov_sparse = np.asarray(gaussian_cl_covariance(ell, probes, cl, noise, scn["f_sky"], True))
    out = dict(cl=cl, noise=noise, cov_sparse=cov_sparse, ell=ell,
               spec=np.array(json.dumps(scn)))
    if cl.shape[0] <= 10:
        out["cov_dense"] = np.asarray(
            gaussian_cl_covariance(ell, probes, cl, noise, scn["f_sky"], False))
    print("%-24s %s  %.1fs" % (scn["name"], cl.shape, time.time() - t), flush=True)
    return out


def run_stages(name, cdict):
    """Per-stage values of the hot path for one cosmology (SURVEY Appendix B style)."""
    import jax_cosmo.background as bk
    import jax_cosmo.power as pw
    import jax_cosmo.transfer as tk

    cosmo = jc.Cosmology(**cdict)
    a = np.array([1.0 / 11.0, 0.1, 0.2, 0.35, 0.5, 0.7, 0.9, 0.99, 1.0])
    k = np.logspace(-4, 2, 25)
    out = dict(cosmo=sc.cosmo_row(cdict), a=a, k=k)
    out["chi"] = np.asarray(bk.radial_comoving_distance(cosmo, a))
    out["growth"] = np.asarray(bk.growth_factor(cosmo, a))
    out["dchioverda"] = np.asarray(bk.dchioverda(cosmo, a))
    out["Esqr"] = np.asarray(bk.Esqr(cosmo, a))
    out["chitab"] = np.asarray(cosmo._workspace["background.radial_comoving_distance"]["chi"])
    out["gtab"] = np.asarray(cosmo._workspace["background.growth_factor"]["g"])
    out["T_eh"] = np.asarray(tk.Eisenstein_Hu(cosmo, k))
    out["sigmasqr8"] = np.asarray(pw.sigmasqr(cosmo, 8.0, tk.Eisenstein_Hu))
    out["plin"] = np.stack([np.asarray(pw.linear_matter_power(cosmo, k, ai)) for ai in a])
    knl, neff, C = pw._halofit_parameters(cosmo, a, tk.Eisenstein_Hu)
    out["k_nl"], out["n_eff"], out["C_hf"] = map(np.asarray, (knl, neff, C))
    out["pnl"] = np.stack([np.asarray(pw.nonlinear_matter_power(cosmo, k, np.atleast_1d(ai)))
                           for ai in a])
    # radial tracer kernels at a few redshifts, ell=100
    z = np.array([0.0, 0.05, 0.3, 0.8, 1.5, 3.0, 6.0, 9.5, 10.0])
    scn = sc.scenario("x", cdict, [100.0], [sc.sources(5, 2.0, True), sc.lenses(5, 2.0, True)])
    probes = sc.build_probes(scn, jc)
    out["z"] = z
    out["kernel_wl_ext"] = np.asarray(probes[0].kernel(cosmo, z, 100.0))
    out["kernel_nc_ext"] = np.asarray(probes[1].kernel(cosmo, z, 100.0))
    scn = sc.scenario("x", cdict, [100.0], [sc.sources(5, 2.0), sc.lenses(5, 2.0)])
    probes = sc.build_probes(scn, jc)
    out["kernel_wl"] = np.asarray(probes[0].kernel(cosmo, z, 100.0))
    out["kernel_nc"] = np.asarray(probes[1].kernel(cosmo, z, 100.0))
    print("stages %-16s done" % name, flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--stages", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if args.stages or args.only is None:
        row0 = dict(zip(sc.COSMO_KEYS, sc.config5_cosmologies(1)[0]))
        for name, c in [("planck15", sc.PLANCK15), ("testcosmo", sc.TESTCOSMO),
                        ("wcdm", sc.WCDM), ("cfg5row0", row0)]:
            np.savez(os.path.join(OUT, "stages_%s.npz" % name), **run_stages(name, c))
    if not args.stages:
        for scn in sc.golden_scenarios():
            if args.only and scn["name"] != args.only:
                continue
            np.savez(os.path.join(OUT, "cl_%s.npz" % scn["name"]), **run_scenario(scn))
