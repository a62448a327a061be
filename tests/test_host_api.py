"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
reference-style objects flatten into the POD descriptor, unsupported options raise (no fallback)."""
import os
import re

import numpy as np
import pytest
from conftest import ROOT

from oracle import scenarios as sc


def test_library_exports_every_declared_symbol():
    from jax_cosmo_b200 import _native
    lib = _native.load_library()
    header = open(os.path.join(ROOT, "include", "jc_b200.h")).read()
    declared = set(re.findall(r"\b(jc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.jc_abi_version() == _native.JC_ABI_VERSION
    assert lib.jc_status_string(-2).decode().startswith("configuration not supported")


def test_struct_sizes_match_header():
    """ctypes mirrors of the POD structs have the C layout (8-byte aligned doubles)."""
    import ctypes as C

    from jax_cosmo_b200 import _native
    assert C.sizeof(_native.jc_nz) == 8 + 8 * (4 + 4 + 2) + 4 * 8
    assert C.sizeof(_native.jc_bias) == 8 + 24
    assert C.sizeof(_native.jc_tracer) == 8 + C.sizeof(_native.jc_nz) + C.sizeof(_native.jc_bias) + 24
    assert C.sizeof(_native.jc_problem) == 24 + 32 * C.sizeof(_native.jc_tracer)  # 6 int32 header (ABI 2)
    assert C.sizeof(_native.jc_ws_layout) == 13 * 8


def test_problem_flattening(jc):
    from jax_cosmo_b200 import _native
    scn = [s for s in sc.golden_scenarios() if s["name"] == "cfg2_extended_wcdm"][0]
    probes = sc.build_probes(scn, jc)
    pb = _native.build_problem(probes, *sc.build_fns(scn, jc))
    flat = sc.flatten_spec(scn)
    assert pb.n_tracers == len(flat["tracers"]) == 10
    assert pb.nonlinear == _native.JC_PK_HALOFIT
    for t, ref in enumerate(flat["tracers"]):
        tr = pb.tracers[t]
        assert tr.kind == (_native.JC_TRACER_WL if ref["kind"] == "wl" else _native.JC_TRACER_NC)
        assert list(tr.nz.params)[:3] == ref["nz"]["params"]
        assert list(tr.nz.shifts)[:tr.nz.n_shifts] == ref["nz"]["shifts"]
        assert tr.nz.gals_per_arcmin2 == ref["nz"]["gals_per_arcmin2"]  # shift drops n_gal (redshift.py:16)
        assert tr.nz.zmax == ref["nz"]["zmax"] and tr.probe_zmax == ref["probe_zmax"]
        if ref["kind"] == "wl":
            assert tr.m_bias == ref["m"] and tr.sigma_e == ref["sigma_e"] and tr.ia_enabled == 1
            assert tr.bias.family == _native.JC_BIAS["des_y1_ia"]
        else:
            assert tr.bias.family == _native.JC_BIAS["inverse_growth"]
            assert tr.bias.params[0] == ref["bias"]["params"][0]


def test_orderings_and_noise(jc):
    import jax_cosmo_b200.angular_cl as acl
    from oracle import cl_oracle as o
    scn = sc.golden_scenarios()[0]
    probes = sc.build_probes(scn, jc)
    assert acl._get_cl_ordering(probes) == o.cl_ordering(4)
    blocks = acl._get_cov_blocks_ordering(probes)
    pairs = o.cl_ordering(4)

    def find(a, b):
        return pairs.index((a, b)) if (a, b) in pairs else pairs.index((b, a))

    want = [(find(i, m), find(j, n), find(i, n), find(j, m)) for i, j in pairs for m, n in pairs]
    assert blocks == want
    # noise_cl is host-only bookkeeping (probes.py:210-223,274-281 / angular_cl.py:101-117)
    assert np.array_equal(jc.cl.noise_cl(scn["ell"], probes), o.noise_cl(scn["ell"], sc.flatten_spec(scn)))


def test_unsupported_options_raise(jc):
    from jax_cosmo_b200 import _native
    nz = jc.redshift.smail_nz(1.0, 2.0, 1.0)
    wl = jc.probes.WeakLensing([nz])
    with pytest.raises(NotImplementedError):
        _native.build_problem([wl], transfer_fn=lambda *a: None)
    with pytest.raises(NotImplementedError):
        _native.build_problem([wl], nonlinear_fn=lambda *a: None)
    # variants are selected like in the reference: functools.partial over the module functions
    from functools import partial
    pbs = _native.build_problem([wl], partial(jc.transfer.Eisenstein_Hu, type="eisenhu"),
                                partial(jc.power.halofit, prescription="smith2003"))
    assert pbs.transfer == _native.JC_TF_EH_NOWIGGLE and pbs.nonlinear == _native.JC_PK_HALOFIT_SMITH
    with pytest.raises(NotImplementedError):
        _native.build_problem([wl], nonlinear_fn=partial(jc.power.halofit, prescription="mead2020"))
    with pytest.raises(NotImplementedError):
        _native.build_problem([wl], transfer_fn=partial(jc.transfer.Eisenstein_Hu, type="bbks"))
    # delta_nz: weak lensing without IA only (the reference raises in density_kernel / nla_kernel)
    assert _native.build_problem([jc.probes.WeakLensing([jc.redshift.delta_nz(1.0)])]).tracers[0].nz.family == 3
    with pytest.raises(NotImplementedError):
        _native.build_problem([jc.probes.NumberCounts([jc.redshift.delta_nz(1.0)], jc.bias.constant_linear_bias(1.0))])
    with pytest.raises(NotImplementedError):
        _native.build_problem([jc.probes.WeakLensing([jc.redshift.delta_nz(1.0)], ia_bias=jc.bias.constant_linear_bias(1.0))])
    # kde_nz: arrays travel by pointer, the plan-cache key is built from their contents
    z1, w1 = np.array([0.5, 0.7, 1.1]), np.array([1.0, 0.5, 0.25])
    pa = _native.build_problem([jc.probes.WeakLensing([jc.redshift.kde_nz(z1, w1, bw=0.1)])])
    pb_ = _native.build_problem([jc.probes.WeakLensing([jc.redshift.kde_nz(z1.copy(), w1.copy(), bw=0.1)])])
    pc = _native.build_problem([jc.probes.WeakLensing([jc.redshift.kde_nz(z1 + 0.1, w1, bw=0.1)])])
    assert pa.tracers[0].nz.family == 4 and pa.tracers[0].nz.kde_n == 3 and pa.tracers[0].nz.kde_bw == 0.1
    assert pa._content_key == pb_._content_key != pc._content_key
    # gamma-growth cosmologies carry a 9th column and select JC_GROWTH_GAMMA (core.py:56-60,104-105)
    row9 = jc.Cosmology(0.3, 0.05, 0.7, 0.96, 0.8, 0.0, -1.0, 0.0, gamma=0.55).to_row()
    assert row9.shape == (9,) and row9[8] == 0.55
    assert _native.build_problem([wl], growth=1).growth == 1 and _native.build_problem([wl]).growth == 0
    with pytest.raises(NotImplementedError):  # stand-alone halofit: unknown prescription (power.py:226,244)
        jc.power.halofit(jc.Planck15(), 1.0, 1.0, jc.transfer.Eisenstein_Hu, prescription="mead2020")
    with pytest.raises(ValueError):
        _native.build_problem([jc.probes.WeakLensing([nz, nz], multiplicative_bias=[0.1])])


def test_no_cpu_fallback(jc):
    """Without a CUDA device the product path raises; it never routes through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    nz = jc.redshift.smail_nz(1.0, 2.0, 1.0)
    with pytest.raises(RuntimeError):
        jc.cl.angular_cl(jc.Planck15(), [10.0, 100.0], [jc.probes.WeakLensing([nz])])
    import jax_cosmo_b200
    pkg = os.path.dirname(jax_cosmo_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no CPU fallback", ""), fn


def test_cosmology_mirror(jc):
    c = jc.Planck15(Omega_c=0.3)
    assert np.allclose(c.to_row(), [0.3, 0.0486, 0.6774, 0.9667, 0.8159, 0.0, -1.0, 0.0])
    assert c.Omega_m == 0.3 + 0.0486 and c.Omega_de == 1.0 - c.Omega_m
    params, flags = c.tree_flatten()
    assert len(params) == 8 and flags == {"gamma_growth": False}
    assert np.array_equal(type(c).tree_unflatten(flags, params).to_row(), c.to_row())


def test_sparse_to_dense(jc):
    rng = np.random.default_rng(0)
    s = rng.random((3, 2, 4))
    d = jc.sparse.to_dense(s)
    assert d.shape == (12, 8)
    for i in range(3):
        for j in range(2):
            assert np.array_equal(d[i * 4:(i + 1) * 4, j * 4:(j + 1) * 4], np.diag(s[i, j]))


def test_process_options_round_trip():
    """jc_set_option / jc_get_option (no GPU needed): defaults of the hot path, bounds, unknown names."""
    from jax_cosmo_b200 import _native
    assert _native.get_option("jvp_group") == 4.0       # tangent directions per forward-mode pass
    assert _native.get_option("jvp_adjoint") == 1.0     # reverse-sweep K3 for 3..8 directions
    assert _native.get_option("lens_mma") == 0.0        # the scalar lens kernel is the default
    assert _native.get_option("power_exact") == 0.0 and _native.get_option("contract_eps") == 1e-20
    try:
        for name, good, bad in (("jvp_group", 2, 5), ("jvp_group", 1, 0), ("contract_kernel", 3, 4)):
            _native.set_option(name, good)
            assert _native.get_option(name) == float(good)
            with pytest.raises(ValueError):
                _native.set_option(name, bad)
        _native.set_option("lens_mma", 1)
        assert _native.get_option("lens_mma") == 1.0
        with pytest.raises(ValueError):
            _native.set_option("no_such_option", 1)
        with pytest.raises(ValueError):
            _native.get_option("no_such_option")
    finally:
        _native.set_option("jvp_group", 4)
        _native.set_option("contract_kernel", 0)
        _native.set_option("lens_mma", 0)


def test_hessian_argument_checks(jc):
    """likelihood.gaussian_cl_log_likelihood_hessian validates its arguments before anything touches the GPU."""
    nz = [jc.redshift.smail_nz(1.0, 2.0, 1.0)]
    probes = [jc.probes.WeakLensing(nz)]
    ell = np.logspace(1, 3, 8)
    data = np.zeros(8)
    row = jc.Planck15().to_row()
    with pytest.raises(ValueError, match="one cosmology"):
        jc.likelihood.gaussian_cl_log_likelihood_hessian(np.stack([row, row]), data, ell, probes)
    with pytest.raises(ValueError, match="unknown parameter"):
        jc.likelihood.gaussian_cl_log_likelihood_hessian(row, data, ell, probes, params=("Omega_x",))
    with pytest.raises(ValueError, match="unknown parameter"):
        jc.likelihood.gaussian_cl_log_likelihood_hessian(row, data, ell, probes, params=("gamma",))  # 8-column row: no gamma
    with pytest.raises(ValueError, match="rel_step"):
        jc.likelihood.gaussian_cl_log_likelihood_hessian(row, data, ell, probes, rel_step=0.0)
    # the fixed-covariance form checks shapes before it builds a plan
    cov = np.zeros((1, 1, 8))
    with pytest.raises(ValueError, match="one cosmology"):
        jc.likelihood.gaussian_log_likelihood_hessian(np.stack([row, row]), data, cov, ell, probes)
    with pytest.raises(ValueError, match="sparse"):
        jc.likelihood.gaussian_log_likelihood_hessian(row, data, np.zeros((8, 8)), ell, probes)
    with pytest.raises(ValueError, match="elements"):
        jc.likelihood.gaussian_log_likelihood_hessian(row, data[:-1], cov, ell, probes)
    with pytest.raises(ValueError, match="unknown parameter"):
        jc.likelihood.gaussian_log_likelihood_hessian(row, data, cov, ell, probes, params=("tau",))


def test_direction_order():
    """Forward-mode directions along h, n_s, sigma8 cannot move the tracer kernels: the host wrappers put them last (stable)."""
    from jax_cosmo_b200 import _native
    wcdm = np.eye(8)[[0, 1, 2, 3, 4, 6, 7]]            # Omega_c, Omega_b, h, n_s, sigma8, w0, wa
    assert list(_native.direction_order(wcdm)) == [0, 1, 5, 6, 2, 3, 4]
    mixed = np.zeros((3, 9))
    mixed[0, 4] = 1.0                                   # sigma8
    mixed[1, [2, 8]] = 1.0, 0.5                         # h + gamma/2: moves R through the growth factor
    mixed[2, 3] = 1.0                                   # n_s
    assert list(_native.direction_order(mixed)) == [1, 0, 2]
    assert list(_native.direction_order(np.eye(8)[[2, 3]])) == [0, 1]
