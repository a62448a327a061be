"""Scenario specs shared by the golden generator, the oracle and the parity tests.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  A scenario is a plain dict so
that the *same* spec can be turned into

  * reference objects   (``build_probes(spec, jax_cosmo)`` -- run on the NumPy jax shim),
  * product objects     (``build_probes(spec, jax_cosmo_b200)`` -- identical constructor names),
  * the oracle's flat problem description (``flatten_spec``).

Vocabulary and defaults follow the reference: ``smail_nz(a, b, z0, gals_per_arcmin2, zmax)``
(redshift.py:15-77), ``systematic_shift(parent, bias)`` (redshift.py:159-171),
``WeakLensing(bins, ia_bias, multiplicative_bias, sigma_e)`` (probes.py:150-168),
``NumberCounts(bins, bias)`` (probes.py:239-242), bias classes (bias.py:10-57).
Synthetic inputs follow SURVEY.md section 8(d).
"""
import numpy as np

PLANCK15 = dict(Omega_c=0.2589, Omega_b=0.0486, h=0.6774, n_s=0.9667, sigma8=0.8159,
                Omega_k=0.0, w0=-1.0, wa=0.0)  # parameters.py:10-20
TESTCOSMO = dict(Omega_c=0.3, Omega_b=0.05, h=0.7, n_s=0.96, sigma8=0.8,
                 Omega_k=0.0, w0=-1.0, wa=0.0)  # tests/test_angular_cl.py:27-36
WCDM = dict(Omega_c=0.27, Omega_b=0.045, h=0.72, n_s=0.95, sigma8=0.78,
            Omega_k=0.0, w0=-0.9, wa=0.15)

COSMO_KEYS = ("Omega_c", "Omega_b", "h", "n_s", "sigma8", "Omega_k", "w0", "wa")  # core.py:99-108


def cosmo_row(c):
    """dict -> [8] array in the reference's tree_flatten order (core.py:99-108); [9] with the growth
    index appended for a gamma-growth cosmology (core.py:104-105)."""
    row = [c[k] for k in COSMO_KEYS]
    if c.get("gamma") is not None:
        row.append(c["gamma"])
    return np.array(row, dtype=np.float64)


def config5_cosmologies(B, seed=20240607):
    """i.i.d. uniform wCDM box, columns (Oc,Ob,h,ns,s8,w0,wa) then Omega_k=0 inserted."""
    lo = np.array([0.20, 0.04, 0.60, 0.92, 0.70, -1.3, -0.5])
    hi = np.array([0.35, 0.06, 0.80, 1.00, 0.90, -0.7, 0.5])
    X = lo + (hi - lo) * np.random.default_rng(seed).random((B, 7))
    out = np.zeros((B, 8))
    out[:, :5] = X[:, :5]
    out[:, 6:] = X[:, 5:]
    return out


def smail(a, b, z0, n=1.0, zmax=10.0, shift=None):
    return dict(family="smail", params=[float(a), float(b), float(z0)],
                gals_per_arcmin2=float(n), zmax=float(zmax), shift=shift)


def fu(a, b, c, n=1.0, zmax=10.0, shift=None):
    """fu_nz (redshift.py:80-105)."""
    return dict(family="fu", params=[float(a), float(b), float(c)], gals_per_arcmin2=float(n),
                zmax=float(zmax), shift=shift)


def delta(z0, n=1.0, zmax=10.0):
    """delta_nz source plane (redshift.py:108-123)."""
    return dict(family="delta", params=[float(z0)], gals_per_arcmin2=float(n), zmax=float(zmax), shift=None)


def kde(zcat, weights, bw, n=1.0, zmax=10.0, shift=None):
    """kde_nz (redshift.py:126-156): Gaussian KDE of a catalogue."""
    return dict(family="kde", params=[], zcat=[float(v) for v in zcat], weights=[float(v) for v in weights],
                bw=float(bw), gals_per_arcmin2=float(n), zmax=float(zmax), shift=shift)


def bias(family, *params):
    return dict(family=family, params=[float(p) for p in params])


def wl(bins, ia=None, m=0.0, sigma_e=0.26):
    return dict(kind="wl", bins=bins, ia=ia, m=m, sigma_e=sigma_e)


def nc(bins, b):
    return dict(kind="nc", bins=bins, bias=b)


def sources(nb, n, extended=False):
    z0s = {4: [0.5, 0.75, 1.0, 1.25], 5: [0.3 + 0.2 * i for i in range(5)],
           10: [0.25 + 0.1 * i for i in range(10)]}[nb]
    bins = [smail(1.0, 2.0, z0, n, shift=(0.01 * (-1) ** i if extended else None))
            for i, z0 in enumerate(z0s)]
    if extended:
        return wl(bins, ia=bias("des_y1_ia", 0.5, 0.0, 0.62),
                  m=[0.01 * (-1) ** i for i in range(nb)])
    return wl(bins)


def lenses(nb, n, extended=False):
    z0s = {5: [0.3 + 0.2 * i for i in range(5)], 10: [0.25 + 0.1 * i for i in range(10)]}[nb]
    bins = [smail(2.0, 4.0, z0, n) for z0 in z0s]
    fam = "inverse_growth" if extended else "constant"
    return nc(bins, [bias(fam, 1.0 + 0.1 * i) for i in range(nb)])


ELL_CFG1 = np.logspace(1, 3, 50)
ELL_CFG2 = np.logspace(1, np.log10(3000), 100)


def scenario(name, cosmo, ell, probes, nonlinear="halofit", f_sky=0.25, transfer="eisenhu_osc",
             prescription="takahashi2012"):
    """nonlinear: "halofit" | "linear"; prescription (halofit only): "takahashi2012" | "smith2003"
    (power.py:144); transfer: "eisenhu_osc" | "eisenhu" (transfer.py:10)."""
    return dict(name=name, cosmo=dict(cosmo), ell=[float(x) for x in np.atleast_1d(ell)],
                probes=probes, nonlinear=nonlinear, f_sky=f_sky, transfer=transfer, prescription=prescription)


def golden_scenarios():
    """Scenarios executed through the reference source (slow: a few seconds per ell)."""
    nz1, nz2 = smail(1.0, 2.0, 1.0), smail(1.0, 2.0, 0.5)
    S = []
    # SURVEY Appendix B
    appb = [wl([nz1, nz2]), nc([nz1, nz2], bias("constant", 1.0))]
    S.append(scenario("appB_halofit", PLANCK15, [10.0, 100.0, 1000.0], appb))
    S.append(scenario("appB_linear", PLANCK15, [10.0, 100.0, 1000.0], appb, "linear"))
    # BASELINE config 1: WL 4 Smail bins, linear, subset of logspace(1,3,50)
    S.append(scenario("cfg1_wl4_linear", PLANCK15, ELL_CFG1[[0, 12, 25, 37, 49]],
                      [sources(4, 6.5)], "linear"))
    # BASELINE config 2: 5+5, halofit, subset of the 100 ells
    S.append(scenario("cfg2_3x2pt_5p5", PLANCK15, ELL_CFG2[[0, 33, 66, 99]],
                      [sources(5, 2.0), lenses(5, 2.0)]))
    # extended variant (photo-z shift, NLA IA, m-bias list, inverse-growth bias) on wCDM
    S.append(scenario("cfg2_extended_wcdm", WCDM, ELL_CFG2[[0, 50, 99]],
                      [sources(5, 2.0, True), lenses(5, 2.0, True)]))
    # the reference's own test scenarios (tests/test_angular_cl.py:14-51,108-198), low ell incl.
    ell_t = np.logspace(0.1, 4, 50)[[0, 10, 25, 40, 49]]
    S.append(scenario("reftest_lensing", TESTCOSMO, ell_t, [wl([nz1])]))
    S.append(scenario("reftest_lensing_ia", TESTCOSMO, ell_t,
                      [wl([nz1], ia=bias("inverse_growth", 10.0))]))
    S.append(scenario("reftest_clustering", TESTCOSMO, ell_t, [nc([nz1], bias("constant", 1.0))]))
    # scalar (non-list) multiplicative bias + sigma_e list; per-bin scalar bias on lenses
    S.append(scenario("scalar_options", WCDM, [30.0, 300.0],
                      [wl([nz1, nz2], m=0.02, sigma_e=[0.26, 0.3]),
                       nc([nz2], bias("constant", 1.3))]))
    # config 3/5 shape: 10+10 bins on two random wCDM rows of the config-5 box
    rows = config5_cosmologies(2)
    for r, row in enumerate(rows):
        S.append(scenario("cfg5_10p10_row%d" % r, dict(zip(COSMO_KEYS, row)), ELL_CFG2[[5, 95]],
                          [sources(10, 1.0), lenses(10, 1.0)]))
    # the remaining n(z) families.  delta_nz source planes as in the reference's
    # tests/test_angular_cl.py:54-105 (alone, and mixed with Smail bins), fu_nz with the Martinet+2021
    # parameters, kde_nz of a small seeded catalogue (also under a photo-z shift), as sources and lenses.
    nzs2 = smail(1.4, 2.0, 1.0)
    S.append(scenario("reftest_lensing_delta", TESTCOSMO, ell_t, [wl([delta(1.0)])]))
    S.append(scenario("reftest_lensing_delta_mix", TESTCOSMO, ell_t[[1, 3]], [wl([delta(1.0), nz1, nzs2])]))
    rng = np.random.default_rng(11)
    zcat = np.round(rng.gamma(4.0, 0.2, 48), 6)
    wcat = np.round(rng.uniform(0.2, 1.0, 48), 6)
    S.append(scenario("families_fu_kde", WCDM, [20.0, 200.0, 2000.0],
                      [wl([fu(0.4710, 5.1843, 0.7259, 30.0), kde(zcat, wcat, 0.1, 4.0), kde(zcat, wcat, 0.15, shift=0.02),
                           delta(0.8)], m=[0.0, 0.01, 0.0, -0.01]),
                       nc([fu(0.4710, 5.1843, 0.7259, 10.0), kde(zcat, wcat, 0.1, 3.0)],
                          [bias("constant", 1.1), bias("inverse_growth", 1.3)])]))
    # non-default physics switches: smith2003 halofit (power.py:182-198,239-242) and the
    # no-wiggle Eisenstein-Hu fit (transfer.py:99-105); an open (Omega_k != 0) wCDM model so that the smith2003
    # f1,f2,f3 interpolation (frac != 1) is exercised
    open_wcdm = dict(WCDM, Omega_k=0.04)
    sw = [wl([nz1, nz2], ia=bias("des_y1_ia", 0.5, 0.0, 0.62)), nc([nz2], bias("constant", 1.2))]
    S.append(scenario("switch_smith2003", open_wcdm, [20.0, 200.0, 2000.0], sw, prescription="smith2003"))
    S.append(scenario("switch_nowiggle", WCDM, [20.0, 200.0, 2000.0], sw, transfer="eisenhu"))
    S.append(scenario("switch_nowiggle_smith_linear", open_wcdm, [50.0, 500.0], sw, "linear", transfer="eisenhu"))
    # gamma-parametrised growth, Cosmology(..., gamma=...) (core.py:56-60, background.py:515-582): D(a) enters
    # P_lin, halofit, the NLA kernel and inverse_growth_linear_bias
    swg = [wl([nz1, nz2], ia=bias("des_y1_ia", 0.5, 0.0, 0.62)), nc([nz2, nz1], [bias("constant", 1.2), bias("inverse_growth", 1.1)])]
    S.append(scenario("switch_gamma_growth", dict(WCDM, gamma=0.55), [20.0, 200.0, 2000.0], swg))
    S.append(scenario("switch_gamma_growth_open_linear", dict(open_wcdm, gamma=0.68), [50.0, 500.0], swg, "linear"))
    return S


# ----------------------------------------------------------------------------------------
# spec -> objects of a jax_cosmo-compatible namespace (reference OR product)
# ----------------------------------------------------------------------------------------
def build_nz(spec, ns):
    kw = dict(gals_per_arcmin2=spec["gals_per_arcmin2"], zmax=spec["zmax"])
    fam = spec["family"]
    if fam == "smail":
        nz = ns.redshift.smail_nz(*spec["params"], **kw)
    elif fam == "fu":
        nz = ns.redshift.fu_nz(*spec["params"], **kw)
    elif fam == "delta":
        nz = ns.redshift.delta_nz(*spec["params"], **kw)
    elif fam == "kde":
        nz = ns.redshift.kde_nz(np.array(spec["zcat"]), np.array(spec["weights"]), bw=spec["bw"], **kw)
    else:
        raise ValueError(fam)
    if spec.get("shift") is not None:
        nz = ns.redshift.systematic_shift(nz, spec["shift"])
    return nz


def build_bias(spec, ns):
    cls = {"constant": ns.bias.constant_linear_bias,
           "inverse_growth": ns.bias.inverse_growth_linear_bias,
           "des_y1_ia": ns.bias.des_y1_ia_bias}[spec["family"]]
    return cls(*spec["params"])


def build_probes(scn, ns):
    out = []
    for p in scn["probes"]:
        bins = [build_nz(b, ns) for b in p["bins"]]
        if p["kind"] == "wl":
            ia = p["ia"]
            if isinstance(ia, list):
                ia = [build_bias(b, ns) for b in ia]
            elif ia is not None:
                ia = build_bias(ia, ns)
            out.append(ns.probes.WeakLensing(bins, ia_bias=ia, multiplicative_bias=p["m"],
                                             sigma_e=p["sigma_e"]))
        else:
            b = p["bias"]
            b = [build_bias(x, ns) for x in b] if isinstance(b, list) else build_bias(b, ns)
            out.append(ns.probes.NumberCounts(bins, b))
    return out


def build_cosmo(scn, ns):
    return ns.Cosmology(**scn["cosmo"])


def build_fns(scn, ns):
    """(transfer_fn, nonlinear_fn) as a user of the reference passes them: non-default variants are
    functools.partial objects over the module functions."""
    from functools import partial
    nl = {"halofit": ns.power.halofit, "linear": ns.power.linear}[scn["nonlinear"]]
    if scn["nonlinear"] == "halofit" and scn.get("prescription", "takahashi2012") != "takahashi2012":
        nl = partial(ns.power.halofit, prescription=scn["prescription"])
    tf = ns.transfer.Eisenstein_Hu
    if scn.get("transfer", "eisenhu_osc") != "eisenhu_osc":
        tf = partial(ns.transfer.Eisenstein_Hu, type=scn["transfer"])
    return tf, nl


# ----------------------------------------------------------------------------------------
# spec -> flat tracer list used by the oracle (one entry per tracer = redshift bin)
# ----------------------------------------------------------------------------------------
def flatten_spec(scn):
    tracers = []
    for p in scn["probes"]:
        nb = len(p["bins"])
        # systematic_shift carries its own default zmax=10 (redshift.py:16)
        pz = max((10.0 if b.get("shift") is not None else b["zmax"]) for b in p["bins"])
        for i, b in enumerate(p["bins"]):
            nzd = dict(family=b["family"], params=list(b["params"]), zmax=b["zmax"],
                       zcat=b.get("zcat"), weights=b.get("weights"), bw=b.get("bw"),
                       shifts=[] if b.get("shift") is None else [b["shift"]],
                       # systematic_shift does not inherit gals_per_arcmin2 (redshift.py:16,159)
                       gals_per_arcmin2=(b["gals_per_arcmin2"] if b.get("shift") is None else 1.0))
            if b.get("shift") is not None:
                nzd["zmax"] = 10.0  # systematic_shift's own default zmax (redshift.py:16)
            if p["kind"] == "wl":
                ia = p["ia"]
                ia_i = ia[i] if isinstance(ia, list) else ia
                m = p["m"][i] if isinstance(p["m"], list) else p["m"]
                se = p["sigma_e"][i] if isinstance(p["sigma_e"], list) else p["sigma_e"]
                tracers.append(dict(kind="wl", nz=nzd, ia=ia_i, m=float(m), sigma_e=float(se),
                                    probe_zmax=pz))
            else:
                bb = p["bias"][i] if isinstance(p["bias"], list) else p["bias"]
                tracers.append(dict(kind="nc", nz=nzd, bias=bb, probe_zmax=pz))
    zmax = max(t["probe_zmax"] for t in tracers)
    return dict(tracers=tracers, zmax=zmax, nonlinear=(scn["nonlinear"] == "halofit"),
                transfer=scn.get("transfer", "eisenhu_osc"), prescription=scn.get("prescription", "takahashi2012"))
