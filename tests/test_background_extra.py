"""The remaining public functions of background.py (w, f_de, Omega_m_a, Omega_de_a, dchioverda, growth_rate, a_of_chi) and
power.sigmasqr against golden values produced by the reference on the NumPy shim (tests/golden/background_extra.npz,
oracle/make_golden.py --background): flat, open, closed and gamma-growth cosmologies.  The CPU test pins the oracle, the
GPU test checks jc_grid_background_f64 / jc_a_of_chi_f64 / jc_sigmasqr_f64."""
import json
import os
from functools import partial

import numpy as np
import pytest
from conftest import GOLDEN, relerr

from oracle import cl_oracle as o

RTOL = 1e-9
FIELDS = ("Omega_m_a", "Omega_de_a", "dchioverda", "growth_rate")


def _golden():
    g = np.load(os.path.join(GOLDEN, "background_extra.npz"))
    return g, json.loads(str(g["names"]))


def test_oracle_background_extra_vs_reference_golden():
    g, names = _golden()
    a = g["a"]
    for name in names:
        c = o.Cosmo(g[name + "_row"])
        bg = o.Background(c)
        assert np.array_equal(o.w_de(c, a), g[name + "_w"])
        assert np.max(np.abs(o.f_de(c, a) - g[name + "_f_de"])) < 1e-14
        assert relerr(o.Omega_m_a(c, a), g[name + "_Omega_m_a"]) < 1e-14
        assert relerr(o.Omega_de_a(c, a), g[name + "_Omega_de_a"]) < 1e-14
        assert relerr(o.dchioverda(c, a), g[name + "_dchioverda"]) < 1e-14
        assert relerr(bg.growth_rate(a), g[name + "_growth_rate"]) < 1e-13
        assert relerr(bg.a_of_chi(g["chi"]), g[name + "_a_of_chi"]) < 1e-13
        for key, ttype in (("_sigmasqr", "eisenhu_osc"), ("_sigmasqr_nowiggle", "eisenhu")):
            c2 = o.Cosmo(g[name + "_row"])
            c2.transfer_type = ttype
            got = np.array([o.sigmasqr_raw(c2, r) for r in g["R"]])
            assert relerr(got, g[name + key]) < 1e-13


def _ell_factor(ell):
    return np.sqrt((ell - 1) * ell * (ell + 1) * (ell + 2)) / (ell + 0.5) ** 2


def test_oracle_probe_kernel_functions_vs_reference_golden():
    """probes.weak_lensing_kernel / density_kernel / nla_kernel (probes.py:17-129)."""
    from oracle import scenarios as sc
    g, _ = _golden()
    nzs = json.loads(str(g["pk_nz"]))
    z = g["pk_z"]
    bg = o.Background(o.Cosmo(g["pk_row"]))

    def kernels(probe):
        scn = sc.scenario("k", dict(zip(sc.COSMO_KEYS, g["pk_row"])), [50.0], [probe])
        return o.radial_kernels(bg, sc.flatten_spec(scn)["tracers"], z)[0]

    wl = kernels(sc.wl(nzs)) * _ell_factor(50.0)
    assert np.max(np.abs(wl - g["pk_wl"])) < 1e-12 * np.abs(g["pk_wl"]).max()
    dens = kernels(sc.nc(nzs, [sc.bias("constant", 1.3)] * 2))
    assert np.max(np.abs(dens - g["pk_density"])) < 1e-12 * np.abs(g["pk_density"]).max()
    nla = kernels(sc.wl(nzs, ia=sc.bias("des_y1_ia", 0.5, 0.1, 0.62))) * _ell_factor(50.0) - wl
    assert np.max(np.abs(nla - g["pk_nla"])) < 1e-11 * np.abs(g["pk_nla"]).max()


@pytest.mark.gpu
def test_gpu_probe_kernel_functions_and_named_sparse_products(jc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import scenarios as sc
    g, _ = _golden()
    z = g["pk_z"]
    pzs = [sc.build_nz(s, jc) for s in json.loads(str(g["pk_nz"]))]
    row = g["pk_row"]
    cosmo = jc.Cosmology(*row[:8])
    for key, got in (("pk_wl", jc.probes.weak_lensing_kernel(cosmo, pzs, z, 50.0)),
                     ("pk_density", jc.probes.density_kernel(cosmo, pzs, jc.bias.constant_linear_bias(1.3), z, 50.0)),
                     ("pk_nla", jc.probes.nla_kernel(cosmo, pzs, jc.bias.des_y1_ia_bias(0.5, 0.1, 0.62), z, 50.0))):
        assert got.shape == g[key].shape
        assert np.max(np.abs(got - g[key])) < RTOL * np.abs(g[key]).max(), (key, np.max(np.abs(got - g[key])))
    # the named special cases of sparse.dot (sparse.py:141-292) against dense NumPy products
    rng = np.random.default_rng(2)
    S, S2 = rng.standard_normal((3, 4, 5)), rng.standard_normal((4, 2, 5))
    v4, v3 = rng.standard_normal(20), rng.standard_normal(15)
    D4, D3 = rng.standard_normal((20, 6)), rng.standard_normal((7, 15))
    sp = jc.sparse
    dS, dS2 = sp.to_dense(S), sp.to_dense(S2)
    assert np.allclose(sp.sparse_dot_vec(S, v4), dS @ v4, rtol=1e-13, atol=1e-13)
    assert np.allclose(sp.sparse_dot_dense(S, D4), dS @ D4, rtol=1e-13, atol=1e-13)
    assert np.allclose(sp.vec_dot_sparse(v3, S), v3 @ dS, rtol=1e-13, atol=1e-13)
    assert np.allclose(sp.dense_dot_sparse(D3, S), D3 @ dS, rtol=1e-13, atol=1e-13)
    assert np.allclose(sp.to_dense(sp.sparse_dot_sparse(S, S2)), dS @ dS2, rtol=1e-13, atol=1e-13)
    assert np.allclose(sp.dense_dot_sparse_dot_dense(D3, S, D4), D3 @ dS @ D4, rtol=1e-13, atol=1e-12)
    with pytest.raises(ValueError):
        sp.sparse_dot_vec(S, D4)


@pytest.mark.gpu
def test_gpu_background_extra_vs_reference_golden(jc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g, names = _golden()
    a, chi, R = g["a"], g["chi"], g["R"]
    bk, pw = jc.background, jc.power
    for name in names:
        row = g[name + "_row"]
        cosmo = jc.Cosmology(*row[:8], gamma=(row[8] if len(row) == 9 else None))
        for fn in FIELDS:
            got = getattr(bk, fn)(cosmo, a)
            assert got.shape == a.shape
            assert relerr(got, g[name + "_" + fn]) < RTOL, (name, fn, relerr(got, g[name + "_" + fn]))
        assert np.max(np.abs(bk.w(cosmo, a) - g[name + "_w"])) < 1e-14
        assert np.max(np.abs(bk.f_de(cosmo, a) - g[name + "_f_de"])) < 1e-12
        assert isinstance(bk.growth_rate(cosmo, 0.5), float)
        assert abs(bk.growth_rate(cosmo, 0.5) - g[name + "_growth_rate"][5]) < RTOL
        got = bk.a_of_chi(cosmo, chi)
        assert got.shape == chi.shape and relerr(got, g[name + "_a_of_chi"]) < RTOL, (name, relerr(got, g[name + "_a_of_chi"]))
        assert bk.a_of_chi(cosmo, 1234.5).shape == (1,)
        s2 = pw.sigmasqr(cosmo, R, jc.transfer.Eisenstein_Hu)
        assert relerr(s2, g[name + "_sigmasqr"]) < RTOL, (name, relerr(s2, g[name + "_sigmasqr"]))
        s2 = pw.sigmasqr(cosmo, R, partial(jc.transfer.Eisenstein_Hu, type="eisenhu"))
        assert relerr(s2, g[name + "_sigmasqr_nowiggle"]) < RTOL
        assert abs(pw.sigmasqr(cosmo, 8.0, jc.transfer.Eisenstein_Hu) / g[name + "_sigmasqr"][1] - 1) < RTOL
        # sigma8 normalisation of the path: sigma8^2 = pknorm * sigmasqr(8)  (power.py:47)
        plin = pw.linear_matter_power(cosmo, np.array([0.1]), 1.0)
        tk = jc.transfer.Eisenstein_Hu(cosmo, np.array([0.1]))
        norm = cosmo.sigma8 ** 2 / pw.sigmasqr(cosmo, 8.0, jc.transfer.Eisenstein_Hu)
        assert abs(float(plin) / (norm * 0.1 ** cosmo.n_s * float(tk[0]) ** 2) - 1) < 1e-9
    # many cosmologies x many distances in one call, against the oracle; more than 512 scale factors
    from jax_cosmo_b200 import _native
    rng = np.random.default_rng(5)
    rows = np.array([g["planck15_row"], g["open_wcdm_row"], g["closed_wcdm_row"]])
    rows = np.repeat(rows, 3, axis=0) * (1.0 + 0.02 * rng.standard_normal((9, 8)) * (rows.repeat(3, axis=0) != 0))
    chis = np.sort(rng.uniform(0.0, 9500.0, 700))
    plan = _native.get_grid_plan([1.0], [1.0], nonlinear=_native.JC_PK_LINEAR)
    got = plan.a_of_chi(torch.as_tensor(rows, device="cuda"), torch.as_tensor(chis, device="cuda")).cpu().numpy()
    for i in range(len(rows)):
        ref = o.Background(o.Cosmo(rows[i])).a_of_chi(chis)
        assert relerr(got[i], ref) < RTOL
    many = np.linspace(0.05, 1.0, 700)
    fr = bk.growth_rate(jc.Planck15(), many)
    assert relerr(fr, o.Background(o.Cosmo(jc.Planck15().to_row())).growth_rate(many)) < RTOL
    with pytest.raises(NotImplementedError):
        pw.sigmasqr(jc.Planck15(), 8.0, jc.transfer.Eisenstein_Hu, kmax=100.0)
