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


def run_likelihood():
    """The reference's own likelihood test scenario (tests/test_likelihood.py:12-35) and a 2+2 3x2pt case:
    gaussian_log_likelihood on the sparse covariance, with and without the log-determinant."""
    from jax_cosmo.likelihood import gaussian_log_likelihood
    nz1, nz2 = sc.smail(1.0, 2.0, 1.0), sc.smail(1.0, 2.0, 0.5)
    cases = {
        "reftest": sc.scenario("like_reftest", sc.PLANCK15, np.logspace(1, 3, 5),
                               [sc.nc([nz1, nz2], sc.bias("constant", 1.0))]),
        "3x2pt": sc.scenario("like_3x2pt", sc.WCDM, np.logspace(1, 3, 7),
                             [sc.wl([nz1, nz2]), sc.nc([nz1, nz2], sc.bias("constant", 1.2))]),
    }
    out = {}
    for tag, scn in cases.items():
        cosmo = sc.build_cosmo(scn, jc)
        probes = sc.build_probes(scn, jc)
        ell = np.array(scn["ell"])
        mu, cov = gaussian_cl_covariance_and_mean(cosmo, ell, probes, sparse=True)
        mu, cov = np.asarray(mu), np.asarray(cov)
        data = 1.1 * mu
        out[tag + "_spec"] = np.array(json.dumps(scn))
        out[tag + "_mu"], out[tag + "_cov"], out[tag + "_data"] = mu, cov, data
        out[tag + "_loglike_logdet"] = np.asarray(gaussian_log_likelihood(data, mu, cov, include_logdet=True))
        out[tag + "_loglike_nologdet"] = np.asarray(gaussian_log_likelihood(data, mu, cov, include_logdet=False))
        print("likelihood %-8s %s %s" % (tag, out[tag + "_loglike_logdet"], out[tag + "_loglike_nologdet"]), flush=True)
    return out


def run_grid_functions():
    """Stand-alone background / power functions of the reference (background.py, power.py) on (a, k) grids: flat, open,
    closed and gamma-growth cosmologies; linear, halofit takahashi2012 and smith2003, both Eisenstein-Hu fits."""
    from functools import partial

    import jax_cosmo.background as bk
    import jax_cosmo.power as pw
    import jax_cosmo.transfer as tk

    a = np.array([0.02, 0.09090909090909091, 0.15, 0.25, 0.4, 0.5, 0.62, 0.75, 0.9, 0.97, 1.0])
    k = np.logspace(-3.5, 1.5, 24)
    cosmos = {"planck15": sc.PLANCK15, "open_wcdm": dict(sc.WCDM, Omega_k=0.04), "closed_wcdm": dict(sc.WCDM, Omega_k=-0.03),
              "gamma": dict(sc.WCDM, gamma=0.55)}
    out = dict(a=a, k=k, names=np.array(json.dumps(list(cosmos))))
    # redshift_distribution.__call__ (redshift.py:27-31) for every family and a shifted bin
    rng = np.random.default_rng(1)
    zcat, wcat = rng.gamma(3.0, 0.3, size=64), rng.uniform(0.5, 1.5, size=64)
    specs = {"smail": sc.smail(1.0, 2.0, 0.7, 3.0), "smail_zmax4": sc.smail(2.0, 1.5, 0.5, 1.0, zmax=4.0),
             "fu": sc.fu(0.4710, 5.1843, 0.7259, 30.0), "kde": sc.kde(zcat, wcat, 0.1, 4.0),
             "smail_shift": sc.smail(1.0, 2.0, 0.7, 3.0, shift=0.03)}
    zq = np.concatenate([np.linspace(0.0, 4.0, 33), [6.5, 9.99]])
    out["nz_z"] = zq
    out["nz_specs"] = np.array(json.dumps({k_: dict(v, zcat=None if v.get("zcat") is None else list(v["zcat"]),
                                                      weights=None if v.get("weights") is None else list(v["weights"]))
                                           for k_, v in specs.items()}))
    for nm, spec in specs.items():
        out["nz_" + nm] = np.asarray(sc.build_nz(spec, jc)(zq))
    # probe.kernel(cosmo, z, ell) (probes.py:188-208, 260-272): lensing with IA / m-bias / a delta plane, counts with
    # constant and inverse-growth biases, on an open wCDM cosmology
    nz1, nz2 = sc.smail(1.0, 2.0, 1.0, 2.0), sc.smail(1.0, 2.0, 0.5, 3.0, shift=0.02)
    kscn = sc.scenario("kernels", dict(sc.WCDM, Omega_k=0.04), [100.0],
                       [sc.wl([nz1, nz2], ia=sc.bias("des_y1_ia", 0.5, 0.0, 0.62), m=[0.01, -0.02]),
                        sc.wl([nz1, sc.delta(0.8)]),
                        sc.nc([nz1, nz2], [sc.bias("constant", 1.2), sc.bias("inverse_growth", 1.1)])])
    zk = np.concatenate([[0.0, 0.01], np.linspace(0.05, 3.0, 30)])
    out["kern_z"] = zk
    out["kern_spec"] = np.array(json.dumps(kscn))
    kcosmo = sc.build_cosmo(kscn, jc)
    for i, probe in enumerate(sc.build_probes(kscn, jc)):
        out["kern_%d_ell100" % i] = np.asarray(probe.kernel(kcosmo, zk, 100.0))
        out["kern_%d_ell2" % i] = np.asarray(probe.kernel(kcosmo, zk, 2.0))
    for name, cdict in cosmos.items():
        t = time.time()
        cosmo = jc.Cosmology(**cdict)
        out[name + "_row"] = sc.cosmo_row(cdict)
        out[name + "_chi"] = np.asarray(bk.radial_comoving_distance(cosmo, a))
        out[name + "_chi_transverse"] = np.asarray(bk.transverse_comoving_distance(cosmo, a))
        out[name + "_dA"] = np.asarray(bk.angular_diameter_distance(cosmo, a))
        out[name + "_growth"] = np.asarray(bk.growth_factor(cosmo, a))
        out[name + "_H"] = np.asarray(bk.H(cosmo, a))
        out[name + "_Esqr"] = np.asarray(bk.Esqr(cosmo, a))
        kk, aa = k[:, None], a  # the reference's own pattern (angular_cl.py:75-80): k [..., n_a] against a [n_a]
        cosmo = jc.Cosmology(**cdict)
        out[name + "_plin"] = np.asarray(pw.linear_matter_power(cosmo, kk, aa))
        cosmo = jc.Cosmology(**cdict)
        out[name + "_plin_nowiggle"] = np.asarray(pw.linear_matter_power(cosmo, kk, aa, transfer_fn=partial(tk.Eisenstein_Hu, type="eisenhu")))
        cosmo = jc.Cosmology(**cdict)
        out[name + "_pnl"] = np.asarray(pw.nonlinear_matter_power(cosmo, kk, aa))
        cosmo = jc.Cosmology(**cdict)
        out[name + "_pnl_smith"] = np.asarray(pw.nonlinear_matter_power(cosmo, kk, aa, nonlinear_fn=partial(pw.halofit, prescription="smith2003")))
        cosmo = jc.Cosmology(**cdict)
        out[name + "_pnl_a1"] = np.asarray(pw.nonlinear_matter_power(cosmo, k))  # the notebook's call: a = 1
        out[name + "_tk"] = np.asarray(tk.Eisenstein_Hu(cosmo, k))
        out[name + "_tk_nowiggle"] = np.asarray(tk.Eisenstein_Hu(cosmo, k, type="eisenhu"))
        print("grid %-12s %.1fs" % (name, time.time() - t), flush=True)
    return out


def run_background_extra():
    """The remaining public functions of background.py (w, f_de, Omega_m_a, Omega_de_a, dchioverda, growth_rate, a_of_chi)
    and power.sigmasqr, same four cosmologies as run_grid_functions()."""
    from functools import partial

    import jax_cosmo.background as bk
    import jax_cosmo.power as pw
    import jax_cosmo.transfer as tk

    a = np.array([0.02, 0.09090909090909091, 0.15, 0.25, 0.4, 0.5, 0.62, 0.75, 0.9, 0.97, 1.0])
    chi = np.array([0.0, 3.0, 50.0, 400.0, 1234.5, 2500.0, 4000.0, 6000.0, 9000.0])
    R = np.array([1.0, 8.0, 20.0])
    cosmos = {"planck15": sc.PLANCK15, "open_wcdm": dict(sc.WCDM, Omega_k=0.04), "closed_wcdm": dict(sc.WCDM, Omega_k=-0.03),
              "gamma": dict(sc.WCDM, gamma=0.55)}
    out = dict(a=a, chi=chi, R=R, names=np.array(json.dumps(list(cosmos))))
    for name, cdict in cosmos.items():
        cosmo = jc.Cosmology(**cdict)
        out[name + "_row"] = sc.cosmo_row(cdict)
        for fn in ("w", "f_de", "Omega_m_a", "Omega_de_a", "dchioverda", "growth_rate"):
            out[name + "_" + fn] = np.asarray(getattr(bk, fn)(cosmo, a))
        out[name + "_a_of_chi"] = np.asarray(bk.a_of_chi(cosmo, chi))
        out[name + "_sigmasqr"] = np.array([float(pw.sigmasqr(cosmo, r, tk.Eisenstein_Hu)) for r in R])
        out[name + "_sigmasqr_nowiggle"] = np.array([float(pw.sigmasqr(cosmo, r, partial(tk.Eisenstein_Hu, type="eisenhu"))) for r in R])
    # module-level kernel functions of probes.py (17-129) on the open wCDM cosmology of the kernel scenario
    import jax_cosmo.probes as pr
    nz1, nz2 = sc.smail(1.0, 2.0, 1.0, 2.0), sc.smail(1.0, 2.0, 0.5, 3.0, shift=0.02)
    pzs = [sc.build_nz(nz1, jc), sc.build_nz(nz2, jc)]
    kc = dict(sc.WCDM, Omega_k=0.04)
    zk = np.concatenate([[0.0, 0.01], np.linspace(0.05, 3.0, 30)])
    out["pk_z"], out["pk_row"] = zk, sc.cosmo_row(kc)
    out["pk_nz"] = np.array(json.dumps([nz1, nz2]))
    out["pk_wl"] = np.asarray(pr.weak_lensing_kernel(jc.Cosmology(**kc), pzs, zk, 50.0))
    out["pk_density"] = np.asarray(pr.density_kernel(jc.Cosmology(**kc), pzs, jc.bias.constant_linear_bias(1.3), zk, 50.0))
    out["pk_nla"] = np.asarray(pr.nla_kernel(jc.Cosmology(**kc), pzs, jc.bias.des_y1_ia_bias(0.5, 0.1, 0.62), zk, 50.0))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--stages", action="store_true")
    ap.add_argument("--likelihood", action="store_true")
    ap.add_argument("--grid", action="store_true")
    ap.add_argument("--background", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if args.background:
        np.savez(os.path.join(OUT, "background_extra.npz"), **run_background_extra())
        sys.exit(0)
    if args.grid or (args.only is None and not args.stages and not args.likelihood):
        np.savez(os.path.join(OUT, "grid_functions.npz"), **run_grid_functions())
        if args.grid:
            sys.exit(0)
    if args.likelihood or (args.only is None and not args.stages):
        np.savez(os.path.join(OUT, "likelihood.npz"), **run_likelihood())
        if args.likelihood:
            sys.exit(0)
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
