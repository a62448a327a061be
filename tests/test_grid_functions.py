"""Stand-alone background / matter-power functions (jax_cosmo_b200.background, jax_cosmo_b200.power) against golden
values produced by the reference's background.py / power.py on the NumPy shim (tests/golden/grid_functions.npz,
oracle/make_golden.py --grid): flat, open, closed and gamma-growth cosmologies; linear, halofit takahashi2012 /
smith2003, both Eisenstein-Hu fits.  The CPU test pins the oracle, the GPU test checks the CUDA grid plan."""
import json
import os
from functools import partial

import numpy as np
import pytest
from conftest import GOLDEN, relerr

from oracle import cl_oracle as o

RTOL = 1e-9


def _golden():
    g = np.load(os.path.join(GOLDEN, "grid_functions.npz"))
    return g, json.loads(str(g["names"]))


def test_oracle_grid_functions_vs_reference_golden():
    g, names = _golden()
    a, k = g["a"], g["k"]
    for name in names:
        row = g[name + "_row"]
        c = o.Cosmo(row)
        bg = o.Background(c)
        chi = bg.chi(a, fast=False)
        assert relerr(chi[:-1], g[name + "_chi"][:-1]) < 1e-12 and abs(chi[-1]) < 1e-9
        assert relerr(bg.growth(a, fast=False), g[name + "_growth"]) < 1e-12
        assert relerr(o.H0 * np.sqrt(o.Esqr(c, a)), g[name + "_H"]) < 1e-13
        assert relerr(o.Esqr(c, a), g[name + "_Esqr"]) < 1e-13
        sk = np.sqrt(abs(c.Omega_k))
        ft = chi if c.Omega_k == 0 else (o.RH / sk * (np.sinh if c.Omega_k > 0 else np.sin)(sk * chi / o.RH))
        assert relerr(ft[:-1], g[name + "_chi_transverse"][:-1]) < 1e-12
        assert relerr((a * ft)[:-1], g[name + "_dA"][:-1]) < 1e-12
        for key, ttype in (("_tk", "eisenhu_osc"), ("_tk_nowiggle", "eisenhu")):
            c2 = o.Cosmo(row)
            c2.transfer_type = ttype
            assert relerr(o.eisenstein_hu(c2, k), g[name + key]) < 1e-12
        kk, aa = np.broadcast_arrays(k[:, None], a[None, :])
        for key, ttype, presc in (("_plin", "eisenhu_osc", None), ("_plin_nowiggle", "eisenhu", None),
                                  ("_pnl", "eisenhu_osc", "takahashi2012"), ("_pnl_smith", "eisenhu_osc", "smith2003")):
            c2 = o.Cosmo(row)
            c2.transfer_type = ttype
            c2.prescription = presc or "takahashi2012"
            pw = o.Power(o.Background(c2))
            got = pw.linear(kk, aa) if presc is None else np.stack([pw.halofit(kk[i], aa[i]) for i in range(len(k))])
            assert relerr(got, g[name + key]) < 1e-11, (name, key, relerr(got, g[name + key]))


def _nz_specs(g):
    specs = json.loads(str(g["nz_specs"]))
    for v in specs.values():
        if v.get("zcat") is not None:
            v["zcat"], v["weights"] = np.array(v["zcat"]), np.array(v["weights"])
    return specs


def test_oracle_nz_call_vs_reference_golden():
    """redshift_distribution.__call__ (redshift.py:27-31): oracle n(z) against the reference objects."""
    g, _ = _golden()
    for name, spec in _nz_specs(g).items():
        nzd = dict(family=spec["family"], params=list(spec["params"]), zmax=10.0 if spec.get("shift") is not None else spec["zmax"],
                   zcat=spec.get("zcat"), weights=spec.get("weights"), bw=spec.get("bw"),
                   shifts=[] if spec.get("shift") is None else [spec["shift"]])
        assert relerr(o.nz_eval(nzd, g["nz_z"]), g["nz_" + name], floor=1e-300) < 1e-12, name


def _ell_factor(ell):
    return np.sqrt((ell - 1) * ell * (ell + 1) * (ell + 2)) / (ell + 0.5) ** 2


def test_oracle_probe_kernels_vs_reference_golden():
    """probe.kernel(cosmo, z, ell) (probes.py:188-208, 260-272) per probe: IA + m-bias + shifted bin, a delta plane,
    number counts with constant / inverse-growth bias."""
    from oracle import scenarios as sc
    g, _ = _golden()
    scn = json.loads(str(g["kern_spec"]))
    z = g["kern_z"]
    for i, probe in enumerate(scn["probes"]):
        one = dict(scn, probes=[probe])
        bg = o.Background(o.Cosmo(sc.cosmo_row(scn["cosmo"])))
        R, is_wl = o.radial_kernels(bg, sc.flatten_spec(one)["tracers"], z)
        for ell in (100.0, 2.0):
            ref = g["kern_%d_ell%d" % (i, int(ell))]
            got = R * (_ell_factor(ell) if is_wl[0] else 1.0)
            assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref)), (i, ell)


@pytest.mark.gpu
def test_gpu_probe_kernels_vs_reference_golden(jc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import scenarios as sc
    g, _ = _golden()
    scn = json.loads(str(g["kern_spec"]))
    z = g["kern_z"]
    cosmo = sc.build_cosmo(scn, jc)
    for i, probe in enumerate(sc.build_probes(scn, jc)):
        for ell in (100.0, 2.0):
            ref = g["kern_%d_ell%d" % (i, int(ell))]
            got = probe.kernel(cosmo, z, ell)
            assert got.shape == ref.shape
            scale = np.abs(ref).max(axis=1, keepdims=True)
            assert np.max(np.abs(got - ref) / scale) < RTOL, (i, ell, np.max(np.abs(got - ref) / scale))
    # more redshifts than one grid plan holds
    zz = np.linspace(0.0, 2.5, 600)
    probe = sc.build_probes(scn, jc)[2]
    k600 = probe.kernel(cosmo, zz, 10.0)
    R, _ = o.radial_kernels(o.Background(o.Cosmo(sc.cosmo_row(scn["cosmo"]))), sc.flatten_spec(dict(scn, probes=[scn["probes"][2]]))["tracers"], zz)
    assert k600.shape == (2, 600) and np.max(np.abs(k600 - R)) < RTOL * np.abs(R).max()


@pytest.mark.gpu
def test_gpu_nz_call_vs_reference_golden(jc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import scenarios as sc
    g, _ = _golden()
    for name, spec in _nz_specs(g).items():
        nz = sc.build_nz(spec, jc)
        got = nz(g["nz_z"])
        assert got.shape == g["nz_z"].shape
        assert np.max(np.abs(got - g["nz_" + name])) < RTOL * np.max(np.abs(g["nz_" + name])), name
        assert isinstance(nz(0.5), float)
    with pytest.raises(NotImplementedError):
        jc.redshift.delta_nz(1.0)(0.5)


@pytest.mark.gpu
def test_gpu_background_and_power_vs_reference_golden(jc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g, names = _golden()
    a, k = g["a"], g["k"]
    for name in names:
        row = g[name + "_row"]
        cosmo = jc.Cosmology(*row[:8], gamma=(row[8] if len(row) == 9 else None))
        bk, pw = jc.background, jc.power
        chi = bk.radial_comoving_distance(cosmo, a)
        assert chi.shape == a.shape and relerr(chi[:-1], g[name + "_chi"][:-1]) < RTOL and abs(chi[-1]) < 1e-9
        assert relerr(bk.transverse_comoving_distance(cosmo, a)[:-1], g[name + "_chi_transverse"][:-1]) < RTOL
        assert relerr(bk.angular_diameter_distance(cosmo, a)[:-1], g[name + "_dA"][:-1]) < RTOL
        assert relerr(bk.growth_factor(cosmo, a), g[name + "_growth"]) < RTOL
        assert relerr(bk.H(cosmo, a), g[name + "_H"]) < RTOL
        assert relerr(bk.Esqr(cosmo, a), g[name + "_Esqr"]) < RTOL
        assert isinstance(bk.growth_factor(cosmo, 0.5), float)  # scalar in -> scalar out
        assert abs(bk.growth_factor(cosmo, 0.5) - g[name + "_growth"][5]) < RTOL
        kk = k[:, None]
        nowig = partial(jc.transfer.Eisenstein_Hu, type="eisenhu")
        smith = partial(pw.halofit, prescription="smith2003")
        got = {"_plin": pw.linear_matter_power(cosmo, kk, a),
               "_plin_nowiggle": pw.linear_matter_power(cosmo, kk, a, transfer_fn=nowig),
               "_pnl": pw.nonlinear_matter_power(cosmo, kk, a),
               "_pnl_smith": pw.nonlinear_matter_power(cosmo, kk, a, nonlinear_fn=smith),
               "_pnl_a1": pw.nonlinear_matter_power(cosmo, k),
               "_tk": jc.transfer.Eisenstein_Hu(cosmo, k),
               "_tk_nowiggle": jc.transfer.Eisenstein_Hu(cosmo, k, type="eisenhu")}
        for key, val in got.items():
            assert val.shape == g[name + key].shape, (key, val.shape)
            e = relerr(val, g[name + key])
            assert e < RTOL, (name, key, e)
        # nonlinear_fn=linear is the linear spectrum; the angular_cl selectors stay recognisable by identity
        assert np.array_equal(pw.nonlinear_matter_power(cosmo, kk, a, nonlinear_fn=pw.linear), got["_plin"])
    # more than 512 distinct scale factors: several grid plans behind one call
    many = np.linspace(0.05, 1.0, 700)
    cosmo = jc.Planck15()
    d = jc.background.growth_factor(cosmo, many)
    ref = o.Background(o.Cosmo(cosmo.to_row())).growth(many)
    assert relerr(d, ref) < RTOL
    # grid plans are rejected by the angular_cl entry points and vice versa
    from jax_cosmo_b200 import _native
    gp = _native.get_grid_plan([0.1, 1.0], [0.5, 1.0])
    with pytest.raises(ValueError):
        gp.angular_cl_device(torch.as_tensor(cosmo.to_row()[None], device="cuda"))
