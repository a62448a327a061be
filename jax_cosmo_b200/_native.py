"""ctypes binding of libjc_b200.so (include/jc_b200.h) + flattening of the reference-style probe
objects into the POD `jc_problem`.

PyTorch is used only for device memory and streams.  There is NO CPU fallback: a missing library
or a missing CUDA device raises.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjc_b200.so")

JC_ABI_VERSION = 2
JC_MAX_TRACERS = 32
JC_MAX_SHIFTS = 4
JC_OK, JC_ERR_INVALID, JC_ERR_UNSUPPORTED, JC_ERR_WORKSPACE, JC_ERR_CUDA, JC_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
JC_NZ = {"smail": 1, "fu": 2, "delta": 3, "kde": 4}
JC_BIAS = {"constant": 1, "inverse_growth": 2, "des_y1_ia": 3}
JC_TRACER_WL, JC_TRACER_NC = 1, 2
JC_PK_LINEAR, JC_PK_HALOFIT, JC_PK_HALOFIT_SMITH = 0, 1, 2
JC_TF_EH_OSC, JC_TF_EH_NOWIGGLE = 1, 2

NODE_FIELDS = ["CHI", "INVCHIC", "LNCHIC", "GEOM", "GROWTH", "HUBBLE", "AMP", "RNL", "LNKNL", "NEFF",
               "CURV", "AN", "BN", "LNCF", "P3", "ALPHA", "BETA", "NU", "E1", "E2", "NQ108", "NSILK", "NAMP", "GK", "MU"]
SCAL_FIELDS = ["LN13KEQ", "INV13KEQ", "BETA_C", "C14_ALPHA_C", "SH_D", "LNKSILK", "ALPHA_B", "BETA_B",
               "BETA_NODE", "FB", "FC", "NS", "PKNORM", "SIGMASQR8", "OMEGA_M", "ALPHA_GAMMA", "OMH_T27"]

# every symbol include/jc_b200.h declares
EXPORTS = ["jc_plan_create", "jc_plan_destroy", "jc_plan_n_tracers", "jc_plan_n_cls", "jc_plan_n_ell", "jc_plan_n_cosmo_params",
           "jc_workspace_bytes", "jc_workspace_layout", "jc_angular_cl_f64", "jc_angular_cl_host_f64",
           "jc_workspace_bytes_jvp", "jc_workspace_bytes_jvp_group", "jc_angular_cl_jvp_f64", "jc_gaussian_loglike_f64", "jc_gaussian_cl_loglike_f64", "jc_fisher_f64", "jc_vjp_f64", "jc_sparse_bmm_f64", "jc_sparse_inv_f64", "jc_debug_stages_f64", "jc_grid_plan_create", "jc_grid_plan_create_probes", "jc_grid_eval_f64", "jc_grid_background_f64", "jc_a_of_chi_f64", "jc_sigmasqr_f64", "jc_nz_eval_f64",
           "jc_noise_f64", "jc_gaussian_cov_f64", "jc_gather_create", "jc_gather_status", "jc_gather_buffer", "jc_gather_connect_ipc",
           "jc_gather_connect_local", "jc_gather_destroy", "jc_angular_cl_gather_f64", "jc_gather_push_f64", "jc_set_option", "jc_get_option", "jc_profile_enable", "jc_profile_read",
           "jc_fp64_peak_tflops", "jc_debug_math_f64", "jc_status_string",
           "jc_last_cuda_error", "jc_abi_version"]


class jc_nz(C.Structure):
    _fields_ = [("family", C.c_int32), ("n_shifts", C.c_int32), ("params", C.c_double * 4),
                ("shifts", C.c_double * JC_MAX_SHIFTS), ("gals_per_arcmin2", C.c_double),
                ("zmax", C.c_double), ("kde_z", C.POINTER(C.c_double)), ("kde_w", C.POINTER(C.c_double)),
                ("kde_n", C.c_int64), ("kde_bw", C.c_double)]


class jc_bias(C.Structure):
    _fields_ = [("family", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * 3)]


class jc_tracer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ia_enabled", C.c_int32), ("nz", jc_nz), ("bias", jc_bias),
                ("m_bias", C.c_double), ("sigma_e", C.c_double), ("probe_zmax", C.c_double)]


class jc_problem(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_tracers", C.c_int32), ("transfer", C.c_int32),
                ("nonlinear", C.c_int32), ("growth", C.c_int32), ("reserved", C.c_int32),
                ("tracers", jc_tracer * JC_MAX_TRACERS)]


class jc_ws_layout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("chunk", "node_stride", "ell_stride", "tracer_stride", "chitab",
                                         "gtab", "scal", "stab", "node", "rker", "vtab", "ellpow", "total")]


_lib = None
_lock = threading.Lock()


def load_library():
    """Load libjc_b200.so; raise loudly when it has not been built (no fallback path exists)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "jax_cosmo_b200: %s not found. Build it with jax_cosmo_b200/csrc/build.sh "
                "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
        lib.jc_plan_create.argtypes = [C.POINTER(jc_problem), dp, i32, i32, C.POINTER(vp)]
        lib.jc_plan_create.restype = C.c_int
        lib.jc_plan_destroy.argtypes = [vp]
        lib.jc_plan_destroy.restype = None
        for f in ("jc_plan_n_tracers", "jc_plan_n_cls", "jc_plan_n_ell", "jc_plan_n_cosmo_params"):
            getattr(lib, f).argtypes = [vp]
            getattr(lib, f).restype = i32
        lib.jc_workspace_bytes.argtypes = [vp, i64, C.POINTER(C.c_size_t)]
        lib.jc_workspace_bytes.restype = C.c_int
        lib.jc_workspace_layout.argtypes = [vp, C.c_size_t, C.POINTER(jc_ws_layout)]
        lib.jc_workspace_layout.restype = C.c_int
        lib.jc_angular_cl_f64.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
        lib.jc_angular_cl_f64.restype = C.c_int
        lib.jc_workspace_bytes_jvp.argtypes = [vp, i64, C.POINTER(C.c_size_t)]
        lib.jc_workspace_bytes_jvp.restype = C.c_int
        lib.jc_workspace_bytes_jvp_group.argtypes = [vp, i64, i32, C.POINTER(C.c_size_t)]
        lib.jc_workspace_bytes_jvp_group.restype = C.c_int
        lib.jc_angular_cl_jvp_f64.argtypes = [vp, vp, vp, i32, i64, vp, vp, vp, C.c_size_t, vp]
        lib.jc_angular_cl_jvp_f64.restype = C.c_int
        lib.jc_gaussian_loglike_f64.argtypes = [vp, i64, vp, vp, i64, i32, i32, i32, vp, vp, vp]
        lib.jc_gaussian_loglike_f64.restype = C.c_int
        lib.jc_gaussian_cl_loglike_f64.argtypes = [vp, vp, vp, i64, vp, i64, C.c_double, i32, vp, vp, vp, vp]
        lib.jc_gaussian_cl_loglike_f64.restype = C.c_int
        lib.jc_fisher_f64.argtypes = [vp, vp, i64, i32, i32, i32, vp, vp, vp]
        lib.jc_fisher_f64.restype = C.c_int
        lib.jc_vjp_f64.argtypes = [vp, vp, i64, i64, i32, i64, vp, vp]
        lib.jc_vjp_f64.restype = C.c_int
        lib.jc_sparse_bmm_f64.argtypes = [vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64, i32, i32, i32, i32, vp]
        lib.jc_sparse_bmm_f64.restype = C.c_int
        lib.jc_sparse_inv_f64.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
        lib.jc_sparse_inv_f64.restype = C.c_int
        lib.jc_grid_plan_create.argtypes = [i32, i32, i32, dp, i32, dp, i32, i32, C.POINTER(C.c_void_p)]
        lib.jc_grid_plan_create.restype = C.c_int
        lib.jc_grid_eval_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
        lib.jc_grid_plan_create_probes.argtypes = [C.POINTER(jc_problem), dp, i32, i32, C.POINTER(C.c_void_p)]
        lib.jc_grid_plan_create_probes.restype = C.c_int
        lib.jc_grid_eval_f64.restype = C.c_int
        lib.jc_grid_background_f64.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
        lib.jc_grid_background_f64.restype = C.c_int
        lib.jc_a_of_chi_f64.argtypes = [vp, vp, i64, vp, i64, vp, vp, C.c_size_t, vp]
        lib.jc_a_of_chi_f64.restype = C.c_int
        lib.jc_sigmasqr_f64.argtypes = [vp, vp, i64, vp, i32, C.c_double, C.c_double, vp, vp, C.c_size_t, vp]
        lib.jc_sigmasqr_f64.restype = C.c_int
        lib.jc_debug_stages_f64.argtypes = [vp, i32, vp, i64, vp, vp, C.c_size_t, vp]
        lib.jc_debug_stages_f64.restype = C.c_int
        lib.jc_nz_eval_f64.argtypes = [C.POINTER(jc_nz), dp, i64, dp]
        lib.jc_nz_eval_f64.restype = C.c_int
        lib.jc_angular_cl_host_f64.argtypes = [vp, vp, i64, vp]
        lib.jc_angular_cl_host_f64.restype = C.c_int
        lib.jc_noise_f64.argtypes = [vp, dp]
        lib.jc_noise_f64.restype = C.c_int
        lib.jc_gaussian_cov_f64.argtypes = [vp, vp, vp, i64, C.c_double, vp, vp]
        lib.jc_gaussian_cov_f64.restype = C.c_int
        lib.jc_gather_create.argtypes = [i32, i32, i32, C.c_size_t, i32, C.POINTER(vp), C.c_char_p]
        lib.jc_gather_status.argtypes = [vp, C.POINTER(i32)]
        lib.jc_gather_status.restype = C.c_int
        lib.jc_gather_create.restype = C.c_int
        lib.jc_gather_buffer.argtypes = [vp]
        lib.jc_gather_buffer.restype = vp
        lib.jc_gather_connect_ipc.argtypes = [vp, C.c_char_p]
        lib.jc_gather_connect_ipc.restype = C.c_int
        lib.jc_gather_connect_local.argtypes = [vp, C.POINTER(vp), C.POINTER(i32)]
        lib.jc_gather_connect_local.restype = C.c_int
        lib.jc_gather_destroy.argtypes = [vp]
        lib.jc_gather_destroy.restype = C.c_int
        lib.jc_angular_cl_gather_f64.argtypes = [vp, vp, vp, i64, i64, i64, i64, i32, vp, C.c_size_t, vp]
        lib.jc_angular_cl_gather_f64.restype = C.c_int
        lib.jc_gather_push_f64.argtypes = [vp, C.c_size_t, i64, i64, vp]
        lib.jc_gather_push_f64.restype = C.c_int
        lib.jc_set_option.argtypes = [C.c_char_p, C.c_double]
        lib.jc_set_option.restype = C.c_int
        lib.jc_get_option.argtypes = [C.c_char_p, dp]
        lib.jc_get_option.restype = C.c_int
        lib.jc_profile_enable.argtypes = [vp, i32]
        lib.jc_profile_enable.restype = C.c_int
        lib.jc_profile_read.argtypes = [vp, dp, C.POINTER(C.c_int64)]
        lib.jc_profile_read.restype = C.c_int
        lib.jc_fp64_peak_tflops.argtypes = [i32, C.c_double, dp]
        lib.jc_fp64_peak_tflops.restype = C.c_int
        lib.jc_debug_math_f64.argtypes = [i32, vp, vp, i64, vp]
        lib.jc_debug_math_f64.restype = C.c_int
        lib.jc_status_string.argtypes = [C.c_int]
        lib.jc_status_string.restype = C.c_char_p
        lib.jc_last_cuda_error.argtypes = []
        lib.jc_last_cuda_error.restype = C.c_char_p
        lib.jc_abi_version.argtypes = []
        lib.jc_abi_version.restype = i32
        if lib.jc_abi_version() != JC_ABI_VERSION:
            raise ImportError("libjc_b200.so ABI version mismatch")
        _lib = lib
        return lib


class JcError(RuntimeError):
    pass


def check(status, what=""):
    if status == JC_OK:
        return
    lib = load_library()
    msg = lib.jc_status_string(status).decode()
    if status == JC_ERR_CUDA:
        msg += ": " + lib.jc_last_cuda_error().decode()
    if status == JC_ERR_UNSUPPORTED:
        raise NotImplementedError("%s: %s" % (what, msg))
    if status == JC_ERR_INVALID:
        raise ValueError("%s: %s" % (what, msg))
    raise JcError("%s: %s" % (what, msg))


# ------------------------------------------------------------------------------------------------
# reference-style objects -> jc_problem
# ------------------------------------------------------------------------------------------------
def _per_bin(value, i, n, what):
    if isinstance(value, (list, tuple)):
        if len(value) != n:
            raise ValueError("%s: expected %d entries, got %d" % (what, n, len(value)))
        return value[i]
    return value


def _fill_bias(dst, b):
    fam = getattr(b, "_family", None)
    if fam not in JC_BIAS:
        raise NotImplementedError("bias %r is not supported by the B200 path" % type(b).__name__)
    dst.family = JC_BIAS[fam]
    for k, v in enumerate(b.params[:3]):
        dst.params[k] = float(v)


def nz_eval(pz, z):
    """redshift_distribution.__call__ on the device (jc_nz_eval_f64): normalised n(z) at host z values."""
    fam, p, shifts = pz._describe()
    if fam not in JC_NZ:
        raise NotImplementedError("n(z) family %s" % fam)
    if len(shifts) > JC_MAX_SHIFTS:
        raise NotImplementedError("more than %d nested systematic_shift" % JC_MAX_SHIFTS)
    d = jc_nz()
    d.family = JC_NZ[fam]
    keep = []
    if fam == "kde":
        zcat, w, bw = p
        keep = [zcat, w]
        d.kde_z = zcat.ctypes.data_as(C.POINTER(C.c_double))
        d.kde_w = w.ctypes.data_as(C.POINTER(C.c_double))
        d.kde_n = len(zcat)
        d.kde_bw = bw
        p = ()
    for k, v in enumerate(p):
        d.params[k] = v
    d.n_shifts = len(shifts)
    for k, v in enumerate(shifts):
        d.shifts[k] = v
    d.gals_per_arcmin2 = float(pz.gals_per_arcmin2)
    d.zmax = float(pz.zmax)
    zz = np.ascontiguousarray(np.atleast_1d(np.asarray(z, dtype=np.float64)))
    out = np.empty(zz.size, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    check(load_library().jc_nz_eval_f64(C.byref(d), zz.reshape(-1).ctypes.data_as(dp), zz.size, out.ctypes.data_as(dp)),
          "jc_nz_eval_f64")
    del keep
    out = out.reshape(zz.shape)
    return float(out[0]) if np.ndim(z) == 0 else out


def build_problem(probes, transfer_fn=None, nonlinear_fn=None, growth=0):
    """Flatten a list of WeakLensing / NumberCounts probes (reference objects of this package)
    into the C descriptor.  Tracer order = probe order, then bin order (angular_cl.py:15-25)."""
    from jax_cosmo_b200 import power as _power
    from jax_cosmo_b200 import transfer as _transfer
    from jax_cosmo_b200.probes import NumberCounts, WeakLensing

    import functools

    def unwrap(fn, allowed_kw):
        """The reference selects variants with functools.partial(fn, kw=...) (transfer.py:10, power.py:144)."""
        kw = {}
        while isinstance(fn, functools.partial):
            if fn.args:
                raise NotImplementedError("positional arguments bound into transfer_fn / nonlinear_fn")
            kw = dict(fn.keywords, **kw)
            fn = fn.func
        extra = set(kw) - set(allowed_kw)
        if extra:
            raise NotImplementedError("unsupported keyword(s) %s" % sorted(extra))
        return fn, kw

    if transfer_fn is None:
        transfer_fn = _transfer.Eisenstein_Hu
    if nonlinear_fn is None:
        nonlinear_fn = _power.halofit
    tf, tkw = unwrap(transfer_fn, ("type",))
    if tf is not _transfer.Eisenstein_Hu:
        raise NotImplementedError("transfer_fn: only jax_cosmo_b200.transfer.Eisenstein_Hu is on the B200 path")
    ttype = tkw.get("type", "eisenhu_osc")
    if ttype not in ("eisenhu_osc", "eisenhu"):
        raise NotImplementedError("Eisenstein_Hu type %r (transfer.py:155)" % (ttype,))
    nlf, nkw = unwrap(nonlinear_fn, ("prescription",))
    if nlf is _power.halofit:
        presc = nkw.get("prescription", "takahashi2012")
        if presc not in ("takahashi2012", "smith2003"):
            raise NotImplementedError("halofit prescription %r (power.py:226,244)" % (presc,))
        nl = JC_PK_HALOFIT if presc == "takahashi2012" else JC_PK_HALOFIT_SMITH
    elif nlf is _power.linear and not nkw:
        nl = JC_PK_LINEAR
    else:
        raise NotImplementedError("nonlinear_fn: only power.halofit (takahashi2012 / smith2003) and power.linear are on the B200 path")

    pb = jc_problem()
    pb._keepalive = []       # arrays referenced by pointer fields
    pb._content_key = b""    # their contents (the plan cache must not key on addresses)
    pb.abi_version = JC_ABI_VERSION
    pb.transfer = JC_TF_EH_OSC if ttype == "eisenhu_osc" else JC_TF_EH_NOWIGGLE
    pb.nonlinear = nl
    pb.growth = int(growth)  # JC_GROWTH_ODE = 0 / JC_GROWTH_GAMMA = 1 (cosmology rows [B, 9])
    t = 0
    for probe in probes:
        if isinstance(probe, WeakLensing):
            kind = JC_TRACER_WL
        elif isinstance(probe, NumberCounts):
            kind = JC_TRACER_NC
        else:
            raise NotImplementedError("probe %r is not supported by the B200 path" % type(probe).__name__)
        pzs = probe.params[0]
        nb = len(pzs)
        pzmax = float(probe.zmax)
        for i, pz in enumerate(pzs):
            if t >= JC_MAX_TRACERS:
                raise NotImplementedError("more than %d tracers" % JC_MAX_TRACERS)
            tr = pb.tracers[t]
            tr.kind = kind
            fam, p, shifts = pz._describe()
            if fam not in JC_NZ:
                raise NotImplementedError("n(z) family %s" % fam)
            if len(shifts) > JC_MAX_SHIFTS:
                raise NotImplementedError("more than %d nested systematic_shift" % JC_MAX_SHIFTS)
            if fam == "delta" and (kind != JC_TRACER_WL or probe.config.get("ia_enabled") or shifts):
                # the reference raises in density_kernel / nla_kernel (probes.py:82-85,107-110)
                raise NotImplementedError("delta_nz is only implemented for weak lensing without IA")
            tr.nz.family = JC_NZ[fam]
            if fam == "kde":
                zcat, w, bw = p
                pb._keepalive.extend([zcat, w])  # host arrays are read by jc_plan_create
                pb._content_key += zcat.tobytes() + w.tobytes()
                tr.nz.kde_z = zcat.ctypes.data_as(C.POINTER(C.c_double))
                tr.nz.kde_w = w.ctypes.data_as(C.POINTER(C.c_double))
                tr.nz.kde_n = len(zcat)
                tr.nz.kde_bw = bw
                p = ()
            for k, v in enumerate(p):
                tr.nz.params[k] = v
            tr.nz.n_shifts = len(shifts)
            for k, v in enumerate(shifts):
                tr.nz.shifts[k] = v
            tr.nz.gals_per_arcmin2 = float(pz.gals_per_arcmin2)
            tr.nz.zmax = float(pz.zmax)
            tr.probe_zmax = pzmax
            if kind == JC_TRACER_WL:
                m = probe.params[1]
                tr.m_bias = float(_per_bin(m, i, nb, "multiplicative_bias"))
                tr.sigma_e = float(_per_bin(probe.config["sigma_e"], i, nb, "sigma_e"))
                if probe.config["ia_enabled"]:
                    tr.ia_enabled = 1
                    _fill_bias(tr.bias, _per_bin(probe.params[2], i, nb, "ia_bias"))
            else:
                _fill_bias(tr.bias, _per_bin(probe.params[1], i, nb, "bias"))
            t += 1
    if t == 0:
        raise ValueError("no tracers")
    pb.n_tracers = t
    return pb


class Plan:
    """Owns a jc_plan (cosmology-independent device tables for one (probes, ell) problem)."""

    def __init__(self, problem, ell, device=None):
        import torch

        lib = load_library()
        if not torch.cuda.is_available():
            raise JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.ell = np.ascontiguousarray(np.atleast_1d(np.asarray(ell, dtype=np.float64)))
        self.problem = problem
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            st = lib.jc_plan_create(C.byref(problem), self.ell.ctypes.data_as(C.POINTER(C.c_double)),
                                    len(self.ell), self.device, C.byref(handle))
        check(st, "jc_plan_create")
        self._h = handle
        self.T = lib.jc_plan_n_tracers(handle)
        self.P = lib.jc_plan_n_cls(handle)
        self.L = lib.jc_plan_n_ell(handle)
        self.ncp = lib.jc_plan_n_cosmo_params(handle)  # 8, or 9 with the growth index gamma
        self._ws = None
        self._noise_dev = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                load_library().jc_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- sizes -----------------------------------------------------------------------------------
    def workspace_bytes(self, n_cosmo):
        out = C.c_size_t()
        check(load_library().jc_workspace_bytes(self._h, int(n_cosmo), C.byref(out)), "jc_workspace_bytes")
        return out.value

    def workspace_layout(self, ws_bytes):
        lo = jc_ws_layout()
        check(load_library().jc_workspace_layout(self._h, int(ws_bytes), C.byref(lo)), "jc_workspace_layout")
        return lo

    def workspace(self, n_cosmo, jvp_entries=None):
        """Scratch for `n_cosmo` cosmologies, cached per CUDA stream: the device entry points are asynchronous on
        torch's current stream, so callers that overlap batches of one problem on several streams (or threads with
        their own current stream) must not share the K1..K4 tables."""
        import torch

        if jvp_entries is None:
            need = self.workspace_bytes(n_cosmo)
        else:
            out = C.c_size_t()
            check(load_library().jc_workspace_bytes_jvp(self._h, int(jvp_entries), C.byref(out)), "jc_workspace_bytes_jvp")
            need = out.value
        if self._ws is None:
            self._ws = {}
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() * 8 < need:
            if ws is None and len(self._ws) >= 8:  # streams come and go: keep the cache bounded
                self._ws.pop(next(iter(self._ws)))
            ws = torch.empty(need // 8, dtype=torch.float64, device="cuda:%d" % self.device)
            self._ws[key] = ws
        return ws

    def _check_dev(self, t, what, shape=None):
        """Real exceptions (not asserts, which -O strips): CUDA float64 contiguous tensor on the plan's device."""
        import torch

        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise ValueError("%s must be a CUDA tensor" % what)
        if t.dtype != torch.float64:
            raise ValueError("%s must be float64 (jax_enable_x64 semantics), got %s" % (what, t.dtype))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % what)
        if t.device.index != self.device:
            raise ValueError("%s lives on cuda:%s, the plan on cuda:%d" % (what, t.device.index, self.device))
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (what, tuple(shape), tuple(t.shape)))

    def _check_rows(self, rows, what="cosmology rows"):
        if rows.ndim != 2 or rows.shape[1] != self.ncp:
            raise ValueError("%s must have shape [n, %d] for this plan (Omega_c, Omega_b, h, n_s, sigma8, Omega_k, "
                             "w0, wa%s), got %s" % (what, self.ncp, ", gamma" if self.ncp == 9 else "", tuple(rows.shape)))

    # -- calls -----------------------------------------------------------------------------------
    def angular_cl_device(self, cosmo_dev, out=None, workspace=None):
        """cosmo_dev: CUDA float64 tensor [B,8] -> CUDA tensor [B,P,L]; async on torch's current stream."""
        import torch

        self._check_dev(cosmo_dev, "cosmology rows")
        self._check_rows(cosmo_dev)
        B = cosmo_dev.shape[0]
        if out is None:
            out = torch.empty((B, self.P, self.L), dtype=torch.float64, device=cosmo_dev.device)
        else:
            self._check_dev(out, "out", (B, self.P, self.L))
        ws = self.workspace(B) if workspace is None else workspace
        if workspace is not None:
            self._check_dev(workspace, "workspace")
        stream = torch.cuda.current_stream(cosmo_dev.device).cuda_stream
        st = load_library().jc_angular_cl_f64(self._h, cosmo_dev.data_ptr(), B, out.data_ptr(), ws.data_ptr(),
                                              ws.numel() * 8, stream)
        check(st, "jc_angular_cl_f64")
        return out

    def angular_cl_jvp_device(self, cosmo_dev, tangents_dev):
        """cosmo_dev [B,8], tangents_dev [K,8] (CUDA float64) -> (cl [B,P,L], dcl [B,K,P,L]) on the device."""
        import torch

        self._check_dev(cosmo_dev, "cosmology rows")
        self._check_dev(tangents_dev, "tangents")
        self._check_rows(cosmo_dev)
        self._check_rows(tangents_dev, "tangents")
        B, K = cosmo_dev.shape[0], tangents_dev.shape[0]
        need = C.c_size_t()
        if B * K <= 512:  # small batches: room for B*K workspace entries lets the library run all K directions in one pass
            check(load_library().jc_workspace_bytes_jvp(self._h, B * K, C.byref(need)), "jc_workspace_bytes_jvp")
        else:  # throughput: value plane + tangent planes of every direction (reverse-sweep K3) or of one tangent group
            check(load_library().jc_workspace_bytes_jvp_group(self._h, B, K, C.byref(need)), "jc_workspace_bytes_jvp_group")
        ws = torch.empty(need.value // 8, dtype=torch.float64, device=cosmo_dev.device)
        cl = torch.empty((B, self.P, self.L), dtype=torch.float64, device=cosmo_dev.device)
        dcl = torch.empty((B, K, self.P, self.L), dtype=torch.float64, device=cosmo_dev.device)
        stream = torch.cuda.current_stream(cosmo_dev.device).cuda_stream
        st = load_library().jc_angular_cl_jvp_f64(self._h, cosmo_dev.data_ptr(), tangents_dev.data_ptr(), K, B,
                                                  cl.data_ptr(), dcl.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream)
        check(st, "jc_angular_cl_jvp_f64")
        return cl, dcl

    def angular_cl_host(self, cosmo_rows, out=None):
        """cosmo_rows: host float64 [B,8] (numpy or CPU tensor) -> host [B,P,L] (same kind)."""
        import torch

        is_t = isinstance(cosmo_rows, torch.Tensor)
        if is_t:
            if cosmo_rows.is_cuda:
                raise ValueError("angular_cl_host takes host rows; use angular_cl_device for CUDA tensors")
            # torch.tensor(python_floats) is float32: the C side reads doubles, so coerce instead of reinterpreting
            rows = cosmo_rows.to(torch.float64).contiguous()
        else:
            rows = np.ascontiguousarray(cosmo_rows, dtype=np.float64)
        self._check_rows(rows)
        B = rows.shape[0]
        shape = (B, self.P, self.L)
        if out is None:
            out = (torch.empty(shape, dtype=torch.float64, pin_memory=True) if is_t
                   else np.empty(shape, dtype=np.float64))
        elif isinstance(out, torch.Tensor):
            if out.is_cuda or out.dtype != torch.float64 or not out.is_contiguous() or tuple(out.shape) != shape:
                raise ValueError("out must be a contiguous float64 CPU tensor of shape %s" % (shape,))
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous
                  and out.flags.writeable and out.shape == shape):
            raise ValueError("out must be a writeable C-contiguous float64 array of shape %s" % (shape,))
        src = rows.data_ptr() if is_t else rows.ctypes.data
        dst = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
        with torch.cuda.device(self.device):
            st = load_library().jc_angular_cl_host_f64(self._h, src, B, dst)
        check(st, "jc_angular_cl_host_f64")
        return out

    STAGES = ["setup", "lens", "finish", "power", "contract"]

    def profile_enable(self, on=True):
        check(load_library().jc_profile_enable(self._h, 1 if on else 0), "jc_profile_enable")

    def profile_read(self):
        """-> ({stage: ms}, {stage: kernel launches}) summed since the last read."""
        ms = (C.c_double * 5)()
        nl = (C.c_int64 * 5)()
        check(load_library().jc_profile_read(self._h, ms, nl), "jc_profile_read")
        return dict(zip(self.STAGES, list(ms))), dict(zip(self.STAGES, list(nl)))

    def noise(self):
        out = np.zeros(self.T, dtype=np.float64)
        check(load_library().jc_noise_f64(self._h, out.ctypes.data_as(C.POINTER(C.c_double))), "jc_noise_f64")
        return out

    def gaussian_cov_device(self, cl_dev, f_sky=0.25, noise=None):
        """cl_dev CUDA [B,P,L] (signal) -> CUDA [B,P,P,L] sparse-block covariance."""
        import torch

        B = cl_dev.shape[0]
        nv = self.noise() if noise is None else np.asarray(noise, dtype=np.float64)
        noise_dev = torch.as_tensor(nv, device=cl_dev.device)
        cov = torch.empty((B, self.P, self.P, self.L), dtype=torch.float64, device=cl_dev.device)
        stream = torch.cuda.current_stream(cl_dev.device).cuda_stream
        st = load_library().jc_gaussian_cov_f64(self._h, cl_dev.data_ptr(), noise_dev.data_ptr(), B, float(f_sky),
                                                cov.data_ptr(), stream)
        check(st, "jc_gaussian_cov_f64")
        return cov


    def gaussian_cl_loglike_device(self, cl_dev, data_dev, f_sky=0.25, include_logdet=True, noise=None, want_cotangent=False):
        """jc_gaussian_cl_loglike_f64: cl_dev CUDA [B, P, L] (signal), data_dev CUDA [P*L] or [B, P*L] -> loglike [B]
        (and, with want_cotangent, d lnL / d cl [B, P, L]) under the Gaussian covariance of the model spectra, without
        forming the covariance.  Asynchronous on torch's current stream."""
        import torch

        self._check_dev(cl_dev, "cl")
        B = cl_dev.shape[0]
        if tuple(cl_dev.shape) != (B, self.P, self.L):
            raise ValueError("cl must have shape [B, %d, %d]" % (self.P, self.L))
        data_dev = data_dev.reshape(-1, self.P * self.L) if data_dev.dim() > 1 else data_dev
        self._check_dev(data_dev, "data")
        if data_dev.dim() == 1 and data_dev.numel() == self.P * self.L:
            stride = 0
        elif data_dev.dim() == 2 and tuple(data_dev.shape) == (B, self.P * self.L):
            stride = self.P * self.L
        else:
            raise ValueError("data must have %d elements, or [B, %d]" % (self.P * self.L, self.P * self.L))
        nv = self.noise() if noise is None else np.asarray(noise, dtype=np.float64)
        if self._noise_dev is None or noise is not None:
            noise_dev = torch.as_tensor(nv, device=cl_dev.device)
            if noise is None:
                self._noise_dev = noise_dev
        else:
            noise_dev = self._noise_dev
        out = torch.empty(B, dtype=torch.float64, device=cl_dev.device)
        dcl = torch.empty((B, self.P, self.L), dtype=torch.float64, device=cl_dev.device) if want_cotangent else None
        scratch = torch.empty((B, self.L, 2), dtype=torch.float64, device=cl_dev.device)
        stream = torch.cuda.current_stream(cl_dev.device).cuda_stream
        for b0 in range(0, B, 32768):  # grid.y limit of the kernel
            b1 = min(B, b0 + 32768)
            st = load_library().jc_gaussian_cl_loglike_f64(
                self._h, cl_dev[b0:b1].data_ptr(), (data_dev[b0:b1] if stride else data_dev).data_ptr(), stride,
                noise_dev.data_ptr(), b1 - b0, float(f_sky), 1 if include_logdet else 0, out[b0:b1].data_ptr(),
                dcl[b0:b1].data_ptr() if want_cotangent else None, scratch[b0:b1].data_ptr(), stream)
            check(st, "jc_gaussian_cl_loglike_f64")
        return (out, dcl) if want_cotangent else out


class GridPlan(Plan):
    """jc_grid_plan_create: the path's setup and power kernels on a caller-chosen (scale factor, wavenumber) grid
    (stand-alone background.* / power.* functions).  n_a <= 512."""

    def __init__(self, k, a, transfer=JC_TF_EH_OSC, nonlinear=JC_PK_HALOFIT, growth=0, device=None, problem=None):
        """problem=None: background / power grid (jc_grid_plan_create); a jc_problem: the radial kernels of its tracers
        at the scale factors `a` (jc_grid_plan_create_probes; `k` is ignored)."""
        import torch

        lib = load_library()
        if not torch.cuda.is_available():
            raise JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.k = np.ascontiguousarray(np.atleast_1d(np.asarray(k, dtype=np.float64)))
        self.a = np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))
        handle = C.c_void_p()
        dp = C.POINTER(C.c_double)
        self.problem = problem
        with torch.cuda.device(self.device):
            if problem is None:
                st = lib.jc_grid_plan_create(int(transfer), int(nonlinear), int(growth), self.k.ctypes.data_as(dp), len(self.k),
                                             self.a.ctypes.data_as(dp), len(self.a), self.device, C.byref(handle))
            else:
                st = lib.jc_grid_plan_create_probes(C.byref(problem), self.a.ctypes.data_as(dp), len(self.a), self.device,
                                                    C.byref(handle))
        check(st, "jc_grid_plan_create")
        self._h = handle
        self.T = lib.jc_plan_n_tracers(handle)
        self.P = lib.jc_plan_n_cls(handle)
        self.L = lib.jc_plan_n_ell(handle)
        self.ncp = lib.jc_plan_n_cosmo_params(handle)
        self._ws = None
        self._noise_dev = None

    def evaluate(self, cosmo_dev, want=("pk", "chi", "chi_transverse", "growth", "hubble")):
        """cosmo_dev CUDA [B, ncp] -> dict of CUDA tensors: pk [B, n_a, n_k], the others [B, n_a]."""
        import torch

        self._check_dev(cosmo_dev, "cosmology rows")
        self._check_rows(cosmo_dev)
        B, na, nk = cosmo_dev.shape[0], len(self.a), len(self.k)
        out = {}
        for name in ("pk", "chi", "chi_transverse", "growth", "hubble", "transfer", "kernels"):
            if name in want:
                shape = {"pk": (B, na, nk), "transfer": (B, nk), "kernels": (B, self.T, na)}.get(name, (B, na))
                out[name] = torch.empty(shape, dtype=torch.float64, device=cosmo_dev.device)
        ptr = lambda n: out[n].data_ptr() if n in out else None
        ws = self.workspace(B)
        st = load_library().jc_grid_eval_f64(self._h, cosmo_dev.data_ptr(), B, ptr("pk"), ptr("chi"), ptr("chi_transverse"),
                                             ptr("growth"), ptr("hubble"), ptr("transfer"), ptr("kernels"), ws.data_ptr(), ws.numel() * 8,
                                             torch.cuda.current_stream(cosmo_dev.device).cuda_stream)
        check(st, "jc_grid_eval_f64")
        return out

    BG_FIELDS = ("growth_rate", "Omega_m_a", "Omega_de_a", "dchioverda", "w", "f_de")  # JC_BG_* of jc_b200.h

    def _cosmo_check(self, cosmo_dev):
        import torch

        self._check_dev(cosmo_dev, "cosmology rows")
        self._check_rows(cosmo_dev)
        return cosmo_dev.shape[0], self.workspace(cosmo_dev.shape[0]), torch.cuda.current_stream(cosmo_dev.device).cuda_stream

    def background(self, cosmo_dev):
        """jc_grid_background_f64: cosmo_dev CUDA [B, ncp] -> CUDA [B, 6, n_a], rows BG_FIELDS."""
        import torch

        B, ws, stream = self._cosmo_check(cosmo_dev)
        aux = torch.empty((B, len(self.BG_FIELDS), len(self.a)), dtype=torch.float64, device=cosmo_dev.device)
        check(load_library().jc_grid_background_f64(self._h, cosmo_dev.data_ptr(), B, aux.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                                    stream), "jc_grid_background_f64")
        return aux

    def a_of_chi(self, cosmo_dev, chi_dev):
        """jc_a_of_chi_f64: chi_dev CUDA [n_chi] (shared) -> CUDA [B, n_chi]."""
        import torch

        B, ws, stream = self._cosmo_check(cosmo_dev)
        self._check_dev(chi_dev, "chi")
        if chi_dev.dim() != 1:
            raise ValueError("chi must be one-dimensional")
        out = torch.empty((B, chi_dev.numel()), dtype=torch.float64, device=cosmo_dev.device)
        check(load_library().jc_a_of_chi_f64(self._h, cosmo_dev.data_ptr(), B, chi_dev.data_ptr(), chi_dev.numel(), out.data_ptr(),
                                             ws.data_ptr(), ws.numel() * 8, stream), "jc_a_of_chi_f64")
        return out

    def sigmasqr(self, cosmo_dev, R_dev, kmin=0.0001, kmax=1000.0):
        """jc_sigmasqr_f64: R_dev CUDA [n_R] (shared) -> CUDA [B, n_R]."""
        import torch

        B, ws, stream = self._cosmo_check(cosmo_dev)
        self._check_dev(R_dev, "R")
        if R_dev.dim() != 1:
            raise ValueError("R must be one-dimensional")
        out = torch.empty((B, R_dev.numel()), dtype=torch.float64, device=cosmo_dev.device)
        check(load_library().jc_sigmasqr_f64(self._h, cosmo_dev.data_ptr(), B, R_dev.data_ptr(), R_dev.numel(), float(kmin), float(kmax),
                                             out.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream), "jc_sigmasqr_f64")
        return out


_grid_cache = {}


def get_grid_plan(k, a, transfer=JC_TF_EH_OSC, nonlinear=JC_PK_HALOFIT, growth=0):
    import torch

    k = np.ascontiguousarray(np.atleast_1d(np.asarray(k, dtype=np.float64)))
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    key = (k.tobytes(), a.tobytes(), int(transfer), int(nonlinear), int(growth), dev)
    plan = _grid_cache.get(key)
    if plan is None:
        if len(_grid_cache) >= 8:
            _grid_cache.pop(next(iter(_grid_cache)))
        plan = GridPlan(k, a, transfer, nonlinear, growth)
        _grid_cache[key] = plan
    return plan


_pinned = {}
_PINNED_CAP_DOUBLES = (128 << 20) // 8  # 128 MB of page-locked staging per device at most


def to_host(t):
    """CUDA tensor -> fresh NumPy array.  Large results (the 35 MB sparse covariance of config 3) go through a cached
    pinned staging buffer: `.cpu()` into pageable memory costs ~15 ms there (allocation + first-touch page faults
    inside the copy), the staged copy ~4 ms."""
    import torch

    n = t.numel()
    if not t.is_cuda or n * t.element_size() < (1 << 20) or t.dtype != torch.float64:
        return t.cpu().numpy()
    # the staging buffer is capped (a dense (P L)^2 covariance is 3.5 GB at the bench configuration: page-locking
    # that much for the life of the process is not acceptable); larger tensors go through it in pieces, two
    # halves alternating so that the device copy of one piece overlaps the host copy of the previous one
    cap = _PINNED_CAP_DOUBLES
    key = t.device.index
    buf = _pinned.get(key)
    want = min(n, cap)
    if buf is None or buf.numel() < want:
        buf = torch.empty(want, dtype=torch.float64, pin_memory=True)
        _pinned[key] = buf
    out = np.empty(tuple(t.shape), dtype=np.float64)
    flat_out, flat_t = out.reshape(-1), t.reshape(-1)
    if n <= buf.numel():
        buf[:n].copy_(flat_t)
        np.copyto(flat_out, buf[:n].numpy())
        return out
    half = buf.numel() // 2
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    pieces = [(o, min(half, n - o)) for o in range(0, n, half)]
    with torch.cuda.device(t.device):
        for i, (o, m) in enumerate(pieces):
            b = i & 1
            buf[b * half:b * half + m].copy_(flat_t[o:o + m], non_blocking=True)
            ev[b].record()
            if i >= 1:
                po, pm = pieces[i - 1]
                ev[1 - b].synchronize()
                np.copyto(flat_out[po:po + pm], buf[(1 - b) * half:(1 - b) * half + pm].numpy())
        po, pm = pieces[-1]
        b = (len(pieces) - 1) & 1
        ev[b].synchronize()
        np.copyto(flat_out[po:po + pm], buf[b * half:b * half + pm].numpy())
    return out


class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


class PeerGather:
    """jc_gather: this rank's full-size result buffer [rows_total, P, L], mapped into every peer (see jc_b200.h).
    `handle` is the CUDA IPC handle to exchange; call connect_ipc(all handles in rank order) or connect_local(...)."""

    def __init__(self, plan, rows_total, rank, world, push_sms=0):
        import torch

        self.plan, self.rank, self.world, self.rows_total = plan, int(rank), int(world), int(rows_total)
        self.push_sms = int(push_sms)
        nbytes = self.rows_total * plan.P * plan.L * 8
        h = C.c_void_p()
        buf = C.create_string_buffer(64)
        check(load_library().jc_gather_create(self.rank, self.world, plan.device, nbytes, self.push_sms, C.byref(h), buf),
              "jc_gather_create")
        self._h = h
        self.handle = bytes(buf.raw)
        self.ptr = load_library().jc_gather_buffer(h)
        self.full = torch.as_tensor(_DevArray(self.ptr, (self.rows_total, plan.P, plan.L), self),
                                    device="cuda:%d" % plan.device)

    def connect_ipc(self, handles):
        if len(handles) != self.world:
            raise ValueError("need one IPC handle per rank")
        check(load_library().jc_gather_connect_ipc(self._h, b"".join(handles)), "jc_gather_connect_ipc")

    def connect_local(self, peers):
        """peers: the PeerGather objects of all ranks (same process, one per device), in rank order."""
        ptrs = (C.c_void_p * self.world)(*[p.ptr for p in peers])
        devs = (C.c_int32 * self.world)(*[p.plan.device for p in peers])
        check(load_library().jc_gather_connect_local(self._h, ptrs, devs), "jc_gather_connect_local")

    def compute_and_push(self, cosmo_dev, row_offset, sub_chunk, push_rows=0, workspace=None, equal_shards=False):
        """K1..K4 of this rank's rows into rows [row_offset, ...) of the local buffer: K1..K3 on chunks of `sub_chunk`
        cosmologies, the contraction per `push_rows` cosmologies, every finished slice pushed to the peers meanwhile.
        Asynchronous on torch's current stream (ordered after the outgoing pushes)."""
        import torch

        n = int(cosmo_dev.shape[0])
        if n:
            self.plan._check_dev(cosmo_dev, "cosmology rows")
            self.plan._check_rows(cosmo_dev)
        ws = self.plan.workspace(max(min(n, int(sub_chunk) if sub_chunk > 0 else n), 1)) if workspace is None else workspace
        stream = torch.cuda.current_stream(self.plan.device).cuda_stream
        check(load_library().jc_angular_cl_gather_f64(self.plan._h, self._h, cosmo_dev.data_ptr() if n else None, n,
                                                      int(row_offset), int(sub_chunk), int(push_rows), 1 if equal_shards else 0, ws.data_ptr(),
                                                      ws.numel() * 8, stream),
              "jc_angular_cl_gather_f64")

    def pusher_aborted(self):
        """True if the pusher kernel of the last step timed out waiting for the compute stream (synchronises)."""
        out = C.c_int32()
        check(load_library().jc_gather_status(self._h, C.byref(out)), "jc_gather_status")
        return bool(out.value)

    def push(self, row_offset, rows):
        import torch

        stream = torch.cuda.current_stream(self.plan.device).cuda_stream
        check(load_library().jc_gather_push_f64(self._h, self.plan.P * self.plan.L * 8, int(row_offset), int(rows), stream),
              "jc_gather_push_f64")

    def close(self):
        if getattr(self, "_h", None):
            self.full = None
            load_library().jc_gather_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plan_cache = {}


def get_plan(probes, ell, transfer_fn=None, nonlinear_fn=None, device=None, growth=0):
    """Plans are cached per (problem bytes, ell bytes, device)."""
    import torch

    pb = build_problem(probes, transfer_fn, nonlinear_fn, growth)
    ell = np.ascontiguousarray(np.atleast_1d(np.asarray(ell, dtype=np.float64)))
    dev = (torch.cuda.current_device() if torch.cuda.is_available() else -1) if device is None else int(device)
    masked = jc_problem.from_buffer_copy(bytes(pb))  # key on contents, never on host addresses
    for t in range(JC_MAX_TRACERS):
        masked.tracers[t].nz.kde_z = None
        masked.tracers[t].nz.kde_w = None
    key = (bytes(masked), pb._content_key, ell.tobytes(), dev)
    plan = _plan_cache.get(key)
    if plan is None:
        if len(_plan_cache) >= 8:
            _plan_cache.pop(next(iter(_plan_cache)))
        plan = Plan(pb, ell, device=None if dev < 0 else dev)
        _plan_cache[key] = plan
    return plan


def gaussian_loglike_device(data_dev, mu_dev, cov_dev, include_logdet=True):
    """CUDA float64 tensors: data [N] or [B,N], mu [B,N], cov [B,P,P,L] (sparse block layout) -> loglike [B]."""
    import torch

    B, P, _, L = cov_dev.shape
    assert mu_dev.shape == (B, P * L) and cov_dev.is_contiguous() and mu_dev.is_contiguous()
    data_dev = data_dev.contiguous()
    stride = 0 if data_dev.dim() == 1 else P * L
    assert data_dev.shape[-1] == P * L
    out = torch.empty(B, dtype=torch.float64, device=cov_dev.device)
    scratch = torch.empty((B, L, 2), dtype=torch.float64, device=cov_dev.device)
    with torch.cuda.device(cov_dev.device):  # a null stream handle means the current device's default stream
        st = load_library().jc_gaussian_loglike_f64(data_dev.data_ptr(), stride, mu_dev.data_ptr(), cov_dev.data_ptr(), B, P, L,
                                                    1 if include_logdet else 0, out.data_ptr(), scratch.data_ptr(),
                                                    torch.cuda.current_stream(cov_dev.device).cuda_stream)
    check(st, "jc_gaussian_loglike_f64")
    return out


def fisher_device(jac_dev, cov_dev):
    """CUDA float64 tensors: jac [B,K,P,L] (or [B,K,P*L]), cov [B,P,P,L] -> Fisher matrices [B,K,K]."""
    import torch

    B, P, _, L = cov_dev.shape
    K = jac_dev.shape[1]
    jac_dev = jac_dev.reshape(B, K, P * L).contiguous()
    out = torch.empty((B, K, K), dtype=torch.float64, device=cov_dev.device)
    scratch = torch.empty((B, L, K * K + 1), dtype=torch.float64, device=cov_dev.device)
    with torch.cuda.device(cov_dev.device):
        st = load_library().jc_fisher_f64(jac_dev.data_ptr(), cov_dev.contiguous().data_ptr(), B, K, P, L, out.data_ptr(),
                                          scratch.data_ptr(), torch.cuda.current_stream(cov_dev.device).cuda_stream)
    check(st, "jc_fisher_f64")
    return out


def vjp_device(jac_dev, cot_dev):
    """CUDA float64 tensors: jac [B,K,...] and a cotangent [B,...] (or [...] shared by the batch) -> J^T g [B,K]."""
    import torch

    B, K = jac_dev.shape[0], jac_dev.shape[1]
    jac_dev = jac_dev.reshape(B, K, -1).contiguous()
    N = jac_dev.shape[2]
    cot_dev = cot_dev.contiguous()
    if cot_dev.numel() == N:
        stride = 0
    elif cot_dev.numel() == B * N:
        stride = N
    else:
        raise ValueError("cotangent has %d elements, expected %d or %d" % (cot_dev.numel(), N, B * N))
    out = torch.empty((B, K), dtype=torch.float64, device=jac_dev.device)
    with torch.cuda.device(jac_dev.device):
        st = load_library().jc_vjp_f64(jac_dev.data_ptr(), cot_dev.data_ptr(), stride, B, K, N, out.data_ptr(),
                                       torch.cuda.current_stream(jac_dev.device).cuda_stream)
    check(st, "jc_vjp_f64")
    return out


def debug_math(fn, x):
    """Evaluate the kernels' own device math (csrc/jc_math.cuh) on a CUDA float64 tensor.
    fn in {"exp", "log", "sin", "rcbrt", "rcp"}."""
    import torch

    code = {"exp": 0, "log": 1, "sin": 2, "rcbrt": 3, "rcp": 4, "exp_t": 5, "log_t": 6}[fn]
    x = x.contiguous()
    y = torch.empty_like(x)
    st = load_library().jc_debug_math_f64(code, x.data_ptr(), y.data_ptr(), x.numel(),
                                          torch.cuda.current_stream(x.device).cuda_stream)
    check(st, "jc_debug_math_f64")
    return y


_R_COLUMNS = (0, 1, 5, 6, 7, 8)  # Omega_c, Omega_b, Omega_k, w0, wa, gamma: the parameters the tracer kernels R_i(a) depend on


def direction_order(tangents):
    """Stable permutation of K forward-mode directions ([K, 8|9] host array) that puts those able to move the tracer kernels
    first.  The throughput path of jc_angular_cl_jvp_f64 runs K1 / K2 per group of four directions; a LATER group made only of
    directions along h, n_s, sigma8 (dR = 0 identically) needs no K2 pass at all.  Callers that build their tangents on the host
    reorder them with this, call the device entry, and undo the permutation on the (host) result."""
    t = np.atleast_2d(np.asarray(tangents, dtype=np.float64))
    cols = [c for c in _R_COLUMNS if c < t.shape[1]]
    moves = np.any(t[:, cols] != 0.0, axis=1)
    return np.argsort(~moves, kind="stable")


def set_option(name, value):
    """jc_set_option: "power_exact" (0 | 1), "contract_eps" (>= 0; read when a plan is created -- cached plans keep theirs),
    "contract_kernel" (0..3), "jvp_group" (1..4 tangent directions per JVP pass)."""
    check(load_library().jc_set_option(name.encode(), float(value)), "jc_set_option(%s)" % name)
    if name == "contract_eps":
        _plan_cache.clear()


def get_option(name):
    out = C.c_double()
    check(load_library().jc_get_option(name.encode(), C.byref(out)), "jc_get_option(%s)" % name)
    return out.value


def fp64_peak_tflops(mode=0, seconds=0.5):
    out = C.c_double()
    check(load_library().jc_fp64_peak_tflops(int(mode), float(seconds), C.byref(out)), "jc_fp64_peak_tflops")
    return out.value
