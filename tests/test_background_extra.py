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
