"""CPU ORACLE -- vectorised NumPy float64 restatement of jax_cosmo's angular-C_ell hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the checker
(or the timed CPU baseline) -- never as a fallback of the product path.

Parity status: PINNED against the reference source itself.  oracle/make_golden.py executes the
unmodified files under /root/reference/jax_cosmo on a NumPy `jax` shim (oracle/jax_shim) and
freezes the outputs as tests/golden/*.npz; tests/test_oracle_golden.py checks every function
below against those fixtures (rtol 1e-12).  The reference's own tests hold no vectors for this
path tighter than rtol 5e-3 (pyccl comparisons), see SURVEY.md 8(c).

Each function cites the reference file:line it restates.  The restatement keeps the reference's
*discretisation and quirks*  and only re-associates arithmetic:
  - chi(a) RK4 on a y-independent ODE == cumulative Simpson with midpoints (background.py:225-233)
  - growth RK4 == ordered product of 2x2 one-step matrices (background.py:463-481)
  - Romberg(divmax=7) == fixed linear functional over 129 nodes (scipy/integrate.py:67-159)
  - halofit sigma^2(R,a) = D(a)^2 S(R) (power.py:96-111)
"""
import numpy as np

# constants.py:9-27, redshift.py:10
C_LIGHT = 299792.458
RH = 2997.92458
H0 = 100.0
TCMB = 2.726
C_1 = 5.0 * 1e-14
RHOCRIT = 2.7750 * 1e11
STERADIAN_TO_ARCMIN2 = 11818102.86004228

N_CHI = 256  # background.py:199
N_GROWTH = 128  # background.py:443
N_LIMBER = 512  # angular_cl.py:96
N_LENS = 256  # probes.py:51
N_NZNORM = 256  # redshift.py:29
N_HF_K = 256  # power.py:111,138
N_HF_R = 256  # power.py:93


# ----------------------------------------------------------------------------------------------
# complex-step support (oracle/derivatives.py::cs_jacobian)
# ----------------------------------------------------------------------------------------------
# The derivative oracle evaluates this module at theta + i h e_k (h = 1e-30): every analytic operation carries the
# directional derivative in the imaginary part to machine precision, and every DECISION -- nearest-node / bracket
# indices, clips, abs, sign -- is taken on the real part with the derivative following the selected branch, which is
# exactly the "frozen-index" derivative jax.jacfwd returns for the reference.  For real input the helpers
# below are numpy's own functions, so the real path (pinned by tests/golden/) is unchanged.
def _clip(x, lo, hi):
    if not np.iscomplexobj(x):
        return np.clip(x, lo, hi)
    xr = np.real(x)
    out = np.array(x, dtype=np.complex128, copy=True)
    if lo is not None:
        out = np.where(xr < np.real(lo), lo, out)
    if hi is not None:
        out = np.where(xr > np.real(hi), hi, out)
    return out


def _abs(x):
    if not np.iscomplexobj(x):
        return np.abs(x)
    return np.where(np.real(x) < 0, -x, x)


def _scalar(v):
    return complex(v) if np.iscomplexobj(v) and np.imag(v) != 0 else float(np.real(v))


# ----------------------------------------------------------------------------------------------
# scipy/ helpers
# ----------------------------------------------------------------------------------------------
def simps_weights(N):
    """Weights (1,4,2,...,4,1) of composite Simpson; multiply by dx/3 (scipy/integrate.py:193-199)."""
    w = np.ones(N + 1)
    w[1:-1:2] = 4.0
    w[2:-1:2] = 2.0
    return w


def romb_weights(divmax=7):
    """Romberg (scipy/integrate.py:134-159) as a fixed linear functional: weights over the
    2**divmax+1 equispaced nodes such that result = (b-a) * sum(w*f)."""
    n = 2 ** divmax + 1
    # trapezoid rule T_i with 2**i intervals, in units of (b-a), as weight vectors over n nodes
    rows = []
    for i in range(divmax + 1):
        step = 2 ** (divmax - i)
        w = np.zeros(n)
        w[::step] = 1.0
        w[0] = w[-1] = 0.5
        rows.append(w / 2 ** i)
    R = rows
    for k in range(1, divmax + 1):
        f = 4.0 ** k
        R = [(f * R[j + 1] - R[j]) / (f - 1.0) for j in range(len(R) - 1)]
    return R[0]


def interp(x, xp, fp):
    """scipy/interpolate.py:12-37 verbatim semantics (nearest node, neighbour by sign rule,
    linear inter/extrapolation anchored at the nearest node).  Brute-force argmin in chunks."""
    x = np.atleast_1d(np.asarray(x))
    shp = x.shape
    x = x.ravel()
    n = len(xp)
    out = np.empty(x.shape, dtype=np.result_type(x, xp, fp, np.float64))
    step = 1 << 15
    xpr = np.real(xp)  # decisions on real parts (complex-step mode); identity for real tables
    for s in range(0, x.size, step):
        xs = x[s:s + step]
        xsr = np.real(xs)
        ind = np.argmin((xsr[:, None] - xpr[None, :]) ** 2, axis=1)
        ind = np.clip(ind, 1, n - 2)
        xi = xp[ind]
        sgn = np.sign(np.clip(xsr, xpr[1], xpr[-2]) - xpr[ind])
        d = np.where(sgn >= 0, 1, -1)  # copysign(1, s) with s integer: s==0 -> +1
        a = (fp[ind + d] - fp[ind]) / (xp[ind + d] - xp[ind])
        b = fp[ind] - a * xp[ind]
        out[s:s + step] = a * xs + b
    return out.reshape(shp)


def interp_index(x, xp):
    """The (ind, d) pair interp() would choose for an INCREASING table, via searchsorted
    (identical to the argmin rule incl. ties: argmin returns the first minimum)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(xp)
    j = np.clip(np.searchsorted(xp, x), 1, n - 1)  # xp[j-1] <= x < xp[j] inside the table
    dl = (x - xp[j - 1]) ** 2
    dr = (x - xp[j]) ** 2
    ind = np.where(dl <= dr, j - 1, j)
    ind = np.clip(ind, 1, n - 2)
    sgn = np.sign(np.clip(x, xp[1], xp[-2]) - xp[ind])
    d = np.where(sgn >= 0, 1, -1)
    return ind, d


def interp_fast(x, xp, fp):
    ind, d = interp_index(x, xp)
    a = (fp[ind + d] - fp[ind]) / (xp[ind + d] - xp[ind])
    return a * x + (fp[ind] - a * xp[ind])


# ----------------------------------------------------------------------------------------------
# background.py
# ----------------------------------------------------------------------------------------------
class Cosmo:
    """[8] row in tree_flatten order (core.py:99-108) + derived Omega_m, Omega_de (core.py:144-162);
    a 9th entry is the growth index gamma of a gamma-growth cosmology (core.py:56-60,104-105)."""

    def __init__(self, row):
        (self.Omega_c, self.Omega_b, self.h, self.n_s, self.sigma8, self.Omega_k, self.w0,
         self.wa) = [_scalar(v) for v in row[:8]]
        self.gamma = _scalar(row[8]) if len(row) > 8 else None
        self.Omega_m = self.Omega_b + self.Omega_c
        self.Omega_de = (1.0 - self.Omega_k) - self.Omega_m


def w_de(c, a):  # background.py:52
    return c.w0 + (1.0 - a) * c.wa


def f_de(c, a):  # background.py:90
    return -3.0 * (1.0 + c.w0 + c.wa) * np.log(a) + 3.0 * c.wa * (a - 1.0)


def Esqr(c, a):  # background.py:122-126
    return c.Omega_m * np.power(a, -3) + c.Omega_k * np.power(a, -2) + c.Omega_de * np.exp(f_de(c, a))


def Omega_m_a(c, a):  # background.py:168
    return c.Omega_m * np.power(a, -3) / Esqr(c, a)


def Omega_de_a(c, a):  # background.py:196
    return c.Omega_de * np.exp(f_de(c, a)) / Esqr(c, a)


def dchioverda(c, a):  # background.py:294
    return RH / (a ** 2 * np.sqrt(Esqr(c, a)))


def chi_table(c):
    """background.py:223-236 + scipy/ode.py:6-22 (RK4 on y-independent rhs)."""
    atab = np.logspace(-3, 0.0, N_CHI)
    x = np.log(atab)
    h = x[1:] - x[:-1]
    xm = x[:-1] + h / 2

    def f(xx):
        xa = np.exp(xx)
        return dchioverda(c, xa) * xa

    k1, k2, k4 = f(x[:-1]), f(xm), f(x[1:])
    inc = 1.0 / 6.0 * h * (k1 + 2 * k2 + 2 * k2 + k4)
    cum = np.concatenate([[0.0], np.cumsum(inc)])
    return atab, cum[-1] - cum


def growth_table_gamma(c):
    """background.py:515-548,582: ln D by RK4 over t = log(atab) of f = Omega_m(a)^gamma (y-independent
    rhs, so each step is h/6 (k1 + 2 k2 + 2 k3 + k4) with k2 = k3), y0 = log(atab[0]); D = exp(y) / exp(y)[-1]."""
    atab = np.logspace(-3, 0.0, N_GROWTH)
    t = np.log(atab)
    h = t[1:] - t[:-1]

    def f(loga):
        return Omega_m_a(c, np.exp(loga)) ** c.gamma

    k1, k2, k4 = f(t[:-1]), f(t[:-1] + h / 2), f(t[1:])
    inc = 1.0 / 6.0 * h * (k1 + 2 * k2 + 2 * k2 + k4)
    y = [t[0]]
    for n in range(N_GROWTH - 1):
        y.append(y[-1] + inc[n])
    g = np.exp(np.array(y))
    return atab, g / g[-1]


def growth_table(c):
    """background.py:461-481: RK4 in a over logspace(-3,0,128) of y=(D,D'), y0=(a0,1)."""
    if getattr(c, "gamma", None) is not None:
        return growth_table_gamma(c)
    atab = np.logspace(-3, 0.0, N_GROWTH)

    def A(x):
        om, ode = Omega_m_a(c, x), Omega_de_a(c, x)
        q = (2.0 - 0.5 * (om + (1.0 + 3.0 * w_de(c, x)) * ode)) / x
        r = 1.5 * om / x / x
        M = np.zeros(x.shape + (2, 2), dtype=np.result_type(r, np.float64))
        M[..., 0, 1] = 1.0
        M[..., 1, 0] = r
        M[..., 1, 1] = -q
        return M

    a0, a1 = atab[:-1], atab[1:]
    h = a1 - a0
    A0, Am, A1 = A(a0), A(a0 + h / 2), A(a1)
    I = np.eye(2)
    hh = h[:, None, None]
    K1 = A0
    K2 = Am @ (I + hh / 2 * K1)
    K3 = Am @ (I + hh / 2 * K2)
    K4 = A1 @ (I + hh * K3)
    M = I + 1.0 / 6.0 * hh * (K1 + 2 * K2 + 2 * K3 + K4)
    y = np.array([atab[0], 1.0])
    D = [y[0]]
    for n in range(N_GROWTH - 1):
        y = M[n] @ y
        D.append(y[0])
    D = np.array(D)
    return atab, D / D[-1]


def growth_rate_table(c):
    """background.py:478-483: ftab = y[:, 1] / y1[-1] * atab / gtab = a D'(a) / D(a) from the same RK4 solution as growth_table()
    (same one-step matrices, second component kept)."""
    atab = np.logspace(-3, 0.0, N_GROWTH)

    def rhs(y, x):  # background.py:465-475
        om, ode = Omega_m_a(c, x), Omega_de_a(c, x)
        q = (2.0 - 0.5 * (om + (1.0 + 3.0 * w_de(c, x)) * ode)) / x
        r = 1.5 * om / x / x
        return np.array([y[1], -q * y[1] + r * y[0]])

    y = np.array([atab[0], 1.0])
    ys = [y]
    for n in range(N_GROWTH - 1):  # scipy/ode.py:6-22
        t0, h = atab[n], atab[n + 1] - atab[n]
        k1 = rhs(y, t0)
        k2 = rhs(y + h * k1 / 2, t0 + h / 2)
        k3 = rhs(y + h * k2 / 2, t0 + h / 2)
        k4 = rhs(y + h * k3, t0 + h)
        y = y + 1.0 / 6.0 * h * (k1 + 2 * k2 + 2 * k3 + k4)
        ys.append(y)
    ys = np.array(ys)
    return atab, ys[:, 1] * atab / ys[:, 0]


class Background:
    """Per-cosmology tables + query functions (the reference's cosmo._workspace, core.py:64)."""

    def __init__(self, c):
        self.c = c
        self.atab, self.chitab = chi_table(c)
        self.ag, self.gtab = growth_table(c)

    def chi(self, a, fast=True):  # background.py:240-242
        f = interp_fast if fast else interp
        return _clip(f(np.atleast_1d(a), self.atab, self.chitab), 0.0, None)

    def growth(self, a, fast=True):  # background.py:488
        f = interp_fast if fast else interp
        return _clip(f(np.atleast_1d(a), self.ag, self.gtab), 0.0, 1.0)

    def growth_rate(self, a):  # background.py:401-440, 491-512, 551-584
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        if getattr(self.c, "gamma", None) is not None:
            return Omega_m_a(self.c, a) ** self.c.gamma
        atab, ftab = growth_rate_table(self.c)
        return interp(a, atab, ftab)

    def a_of_chi(self, chi):  # background.py:245-267: interp() on the DECREASING chi table (its neighbour rule as written)
        return interp(np.atleast_1d(np.asarray(chi)), self.chitab, self.atab)


# ----------------------------------------------------------------------------------------------
# transfer.py / power.py
# ----------------------------------------------------------------------------------------------
def eisenstein_hu(c, k):
    """transfer.py:47-153, type='eisenhu_osc'."""
    T27 = (TCMB / 2.7) ** 2
    h2 = c.h ** 2
    w_m = c.Omega_m * h2
    w_b = c.Omega_b * h2
    fb = c.Omega_b / c.Omega_m
    fc = (c.Omega_m - c.Omega_b) / c.Omega_m
    k_eq = 7.46e-2 * w_m / T27 / c.h
    z_eq = 2.50e4 * w_m / T27 ** 2
    b1 = 0.313 * np.power(w_m, -0.419) * (1.0 + 0.607 * np.power(w_m, 0.674))
    b2 = 0.238 * np.power(w_m, 0.223)
    z_d = 1291.0 * np.power(w_m, 0.251) / (1.0 + 0.659 * np.power(w_m, 0.828)) * (
        1.0 + b1 * np.power(w_b, b2))
    R_d = 31.5 * w_b / T27 ** 2 * (1.0e3 / z_d)
    R_eq = 31.5 * w_b / T27 ** 2 * (1.0e3 / z_eq)
    sh_d = 2.0 / (3.0 * k_eq) * np.sqrt(6.0 / R_eq) * np.log(
        (np.sqrt(1.0 + R_d) + np.sqrt(R_eq + R_d)) / (1.0 + np.sqrt(R_eq)))
    k_silk = 1.6 * np.power(w_b, 0.52) * np.power(w_m, 0.73) * (
        1.0 + np.power(10.4 * w_m, -0.95)) / c.h
    a1 = np.power(46.9 * w_m, 0.670) * (1.0 + np.power(32.1 * w_m, -0.532))
    a2 = np.power(12.0 * w_m, 0.424) * (1.0 + np.power(45.0 * w_m, -0.582))
    alpha_c = np.power(a1, -fb) * np.power(a2, -(fb ** 3))
    b1 = 0.944 / (1.0 + np.power(458.0 * w_m, -0.708))
    if getattr(c, "transfer_type", "eisenhu_osc") == "eisenhu":  # no-wiggle fit, transfer.py:87-105
        alpha_gamma = (1.0 - 0.328 * np.log(431.0 * w_m) * w_b / w_m
                       + 0.38 * np.log(22.3 * w_m) * (c.Omega_b / c.Omega_m) ** 2)
        gamma_eff = c.Omega_m * c.h * (alpha_gamma + (1.0 - alpha_gamma) / (1.0 + (0.43 * k * sh_d) ** 4))
        q = k * np.power(TCMB / 2.7, 2) / gamma_eff
        L = np.log(2.0 * np.exp(1.0) + 1.8 * q)
        C = 14.2 + 731.0 / (1.0 + 62.5 * q)
        return L / (L + C * q * q)
    b2 = np.power(0.395 * w_m, -0.0266)
    beta_c = 1.0 / (1.0 + b1 * (np.power(fc, b2) - 1.0))

    def T_tilde(k1, alpha, beta):
        q = k1 / (13.41 * k_eq)
        L = np.log(np.exp(1.0) + 1.8 * beta * q)
        C = 14.2 / alpha + 386.0 / (1.0 + 69.9 * np.power(q, 1.08))
        return L / (L + C * q * q)

    f = 1.0 / (1.0 + (k * sh_d / 5.4) ** 4)
    Tc = f * T_tilde(k, 1.0, beta_c) + (1.0 - f) * T_tilde(k, alpha_c, beta_c)
    y = (1.0 + z_eq) / (1.0 + z_d)
    x = np.sqrt(1.0 + y)
    G = y * (-6.0 * x + (2.0 + 3.0 * y) * np.log((x + 1.0) / (x - 1.0)))
    alpha_b = 2.07 * k_eq * sh_d * np.power(1.0 + R_d, -0.75) * G
    beta_node = 8.41 * np.power(w_m, 0.435)
    tilde_s = sh_d / np.power(1.0 + (beta_node / (k * sh_d)) ** 3, 1.0 / 3.0)
    beta_b = 0.5 + fb + (3.0 - 2.0 * fb) * np.sqrt((17.2 * w_m) ** 2 + 1.0)
    Tb = (T_tilde(k, 1.0, 1.0) / (1.0 + (k * sh_d / 5.2) ** 2)
          + alpha_b / (1.0 + (beta_b / (k * sh_d)) ** 3) * np.exp(-np.power(k / k_silk, 1.4))
          ) * np.sinc(k * tilde_s / np.pi)
    return fb * Tb + fc * Tc


_ROMB_W = romb_weights(7)


def sigmasqr_raw(c, R=8.0):
    """power.py:56-78 incl. the log10-limits / exp quirk (A.9-2): x in [-4, 3], k = e**x."""
    lo, hi = np.log10(0.0001), np.log10(1000.0)
    xk = np.linspace(lo, hi, 129)
    k = np.exp(xk)
    x = k * R
    w = 3.0 * (np.sin(x) - x * np.cos(x)) / (x * x * x)
    pk = eisenstein_hu(c, k) ** 2 * k ** c.n_s
    y = (hi - lo) * np.sum(_ROMB_W * (k * (k * w) ** 2 * pk))
    return 1.0 / (2.0 * np.pi ** 2.0) * y


class Power:
    """Per-cosmology power-spectrum state: pknorm and the halofit S(R) table."""

    def __init__(self, bg):
        self.bg = bg
        self.c = bg.c
        self.pknorm = self.c.sigma8 ** 2 / sigmasqr_raw(self.c)  # power.py:47
        self._hf = None

    def linear(self, k, a):  # power.py:42-53; k and a broadcast
        g = self.bg.growth(a).reshape(np.shape(a)) if np.ndim(a) else self.bg.growth(a)[0]
        return k ** self.c.n_s * eisenstein_hu(self.c, k) ** 2 * g ** 2 * self.pknorm

    # -- halofit -------------------------------------------------------------------------------
    def _hf_tables(self):
        if self._hf is None:
            logk = np.linspace(np.log(1e-4), np.log(1e4), N_HF_K + 1)  # power.py:111
            k = np.exp(logk)
            wk = simps_weights(N_HF_K) * ((np.log(1e4) - np.log(1e-4)) / N_HF_K) / 3
            g1 = self.bg.growth(1.0)[0]  # linear_matter_power(cosmo, k) at a=1 (power.py:101)
            d2 = k ** self.c.n_s * eisenstein_hu(self.c, k) ** 2 * g1 ** 2 * self.pknorm \
                * k ** 3 / (2.0 * np.pi ** 2)
            logr = np.linspace(np.log(1e-4), np.log(1e1), N_HF_R)  # power.py:93
            r = np.exp(logr)
            S = np.sum((wk * d2)[:, None] * np.exp(-(np.outer(k, r) ** 2)), axis=0)
            self._hf = (k, wk, d2, logr, S)
        return self._hf

    def halofit_parameters(self, a):
        """power.py:86-141 -> k_nl, n_eff, C at each a; also returns the root index `ind`."""
        k, wk, d2, logr, S = self._hf_tables()
        a = np.atleast_1d(a)
        g2 = self.bg.growth(a) ** 2
        sig = g2[:, None] * S[None, :]  # decreasing in r
        n = N_HF_R
        # interp(1.0, xp=sig, fp=logr), scipy/interpolate.py:25-37, per a (quirk A.9-1)
        sigr = np.real(sig)  # root index and neighbour are decided on values (complex-step: real parts)
        ind = np.clip(np.argmin((1.0 - sigr) ** 2, axis=1), 1, n - 2)
        rows = np.arange(len(a))
        xi = sig[rows, ind]
        xq = np.minimum(np.maximum(1.0, sigr[:, 1]), sigr[:, n - 2])  # clip(x, xp[1], xp[-2])
        d = np.where(np.sign(xq - np.real(xi)) >= 0, 1, -1)
        m = (logr[ind + d] - logr[ind]) / (sig[rows, ind + d] - xi)
        root = m * 1.0 + (logr[ind] - m * xi)
        k_nl = 1.0 / _clip(np.exp(root), 1e-6, None)
        y = np.outer(k, 1.0 / k_nl)  # [257, na]
        res = (wk * d2)[:, None] * np.exp(-(y ** 2)) * g2[None, :]
        r0 = np.sum(2 * res * y ** 2, axis=0)
        r1 = np.sum(4 * res * (y ** 2 - y ** 4), axis=0)
        return k_nl, r0 - 3.0, r0 ** 2 + r1, ind

    def halofit_coeffs(self, a):
        """power.py:175-224 (takahashi2012)."""
        c = self.c
        k_nl, n, C, ind = self.halofit_parameters(a)
        om_m = Omega_m_a(c, a)
        om_de = Omega_de_a(c, a)
        w = w_de(c, a)
        co = dict(k_nl=k_nl, n_eff=n, C_hf=C, root_ind=ind, S_tab=self._hf_tables()[4])
        if getattr(c, "prescription", "takahashi2012") == "smith2003":  # power.py:182-198, 239-242
            co["a_n"] = 10 ** (1.4861 + 1.8369 * n + 1.6762 * n ** 2 + 0.7940 * n ** 3 + 0.1670 * n ** 4 - 0.6206 * C)
            co["b_n"] = 10 ** (0.9463 + 0.9466 * n + 0.3084 * n ** 2 - 0.9400 * C)
            co["c_n"] = 10 ** (-0.2807 + 0.6669 * n + 0.3214 * n ** 2 - 0.0793 * C)
            co["gamma_n"] = 0.8649 + 0.2989 * n + 0.1631 * C
            co["alpha_n"] = 1.3884 + 0.3700 * n - 0.1452 * n ** 2
            co["beta_n"] = 0.8291 + 0.9854 * n + 0.3401 * n ** 2
            co["mu_n"] = 10 ** (-3.5442 + 0.1908 * n)
            co["nu_n"] = 10 ** (0.9585 + 1.2857 * n)
            frac = om_de / (1.0 - om_m)
            co["f1"] = frac * om_m ** (-0.0307) + (1 - frac) * om_m ** (-0.0732)
            co["f2"] = frac * om_m ** (-0.0585) + (1 - frac) * om_m ** (-0.1423)
            co["f3"] = frac * om_m ** (0.0743) + (1 - frac) * om_m ** 0.0725
            return co
        co["mu_n"] = np.zeros_like(n)
        co["a_n"] = 10 ** (1.5222 + 2.8553 * n + 2.3706 * n ** 2 + 0.9903 * n ** 3
                           + 0.2250 * n ** 4 - 0.6038 * C + 0.1749 * om_de * (1 + w))
        co["b_n"] = 10 ** (-0.5642 + 0.5864 * n + 0.5716 * n ** 2 - 1.5474 * C
                           + 0.2279 * om_de * (1 + w))
        co["c_n"] = 10 ** (0.3698 + 2.0404 * n + 0.8161 * n ** 2 + 0.5869 * C)
        co["gamma_n"] = 0.1971 - 0.0843 * n + 0.8460 * C
        co["alpha_n"] = _abs(6.0835 + 1.3373 * n - 0.1959 * n ** 2 - 5.5274 * C)
        co["beta_n"] = (2.0379 - 0.7354 * n + 0.3157 * n ** 2 + 1.2490 * n ** 3
                        + 0.3980 * n ** 4 - 0.1682 * C)
        co["nu_n"] = 10 ** (5.2105 + 3.6902 * n)
        co["f1"] = om_m ** (-0.0307)
        co["f2"] = om_m ** (-0.0585)
        co["f3"] = om_m ** (0.0743)
        return co

    def halofit(self, k, a):
        """power.py:144-262: P_nl(k_i, a_i) elementwise (k and a same shape [na])."""
        co = self.halofit_coeffs(a)
        pklin = self.linear(k, a)
        d2l = k ** 3 * pklin / (2.0 * np.pi ** 2)
        y = k / co["k_nl"]
        d2q = d2l * ((1.0 + d2l) ** co["beta_n"] / (1 + co["alpha_n"] * d2l)) * np.exp(
            -(y / 4.0 + y ** 2 / 8.0))
        d2hp = co["a_n"] * y ** (3 * co["f1"]) / (
            1.0 + co["b_n"] * y ** co["f2"] + (co["c_n"] * co["f3"] * y) ** (3.0 - co["gamma_n"]))
        d2h = d2hp / (1.0 + co["mu_n"] / y + co["nu_n"] / y ** 2)
        return 2.0 * np.pi ** 2 / k ** 3 * (d2q + d2h)


# ----------------------------------------------------------------------------------------------
# redshift.py / bias.py / probes.py
# ----------------------------------------------------------------------------------------------
def pz_fn(nz, z):
    """Un-normalised n(z): smail (redshift.py:75-77) under a chain of systematic_shift
    (redshift.py:169-171), outermost shift first."""
    z = np.asarray(z, dtype=np.float64)
    for s in nz["shifts"]:
        z = np.clip(z - s, 0, None)
    fam = nz["family"]
    if fam == "smail":
        a, b, z0 = nz["params"]
        return z ** a * np.exp(-((z / z0) ** b))
    if fam == "fu":  # redshift.py:103-105
        a, b, c = nz["params"]
        return (z ** a + z ** (a * b)) / (z ** b + c)
    if fam == "kde":  # redshift.py:142-156: Gaussian kernel, weights normalised by their sum
        zcat, w, bw = np.asarray(nz["zcat"]), np.asarray(nz["weights"]), nz["bw"]
        k = (1.0 / np.sqrt(2 * np.pi) / bw) * np.exp(-((zcat.reshape((-1,) + (1,) * z.ndim) - z) ** 2) / (bw ** 2 * 2.0))
        return np.tensordot(w, k, axes=(0, 0)) / np.sum(w)
    raise NotImplementedError(fam)


def nz_norm(nz):  # redshift.py:29-30
    zmax = nz["zmax"]
    x = np.linspace(0.0, zmax, N_NZNORM + 1)
    y = pz_fn(nz, x)
    return (zmax - 0.0) / N_NZNORM / 3 * np.sum(y[0:-1:2] + 4 * y[1::2] + y[2::2])


def nz_eval(nz, z):  # redshift.py:27-31
    return pz_fn(nz, z) / nz_norm(nz)


def bias_eval(b, bg, z):  # bias.py:10-57
    fam, p = b["family"], b["params"]
    if fam == "constant":
        return p[0] * np.ones_like(z)
    if fam == "inverse_growth":
        return p[0] / bg.growth(1.0 / (1.0 + z))
    if fam == "des_y1_ia":
        A, eta, z0 = p
        return A * ((1.0 + z) / (1.0 + z0)) ** eta
    raise NotImplementedError(fam)


def lensing_efficiency(bg, nzs, z, zmax):
    """probes.py:27,44-51: simps over z' in linspace(z, zmax, 257) of
    n_i(z') clip(chi'-chi,0)/clip(chi',1), times (1+z) chi.  Returns [len(nzs), len(z)]."""
    chi = bg.chi(1.0 / (1.0 + z))
    zp = np.linspace(z, zmax, N_LENS + 1)  # [257, nz]
    chip = bg.chi(1.0 / (1.0 + zp))
    g = _clip(chip - chi, 0, None) / _clip(chip, 1.0, None)
    dx = (zmax - z) / N_LENS
    w = simps_weights(N_LENS)[:, None]
    out = []
    for nz in nzs:
        out.append(dx / 3 * np.sum(w * nz_eval(nz, zp) * g, axis=0) * (1.0 + z) * chi)
    return np.stack(out, axis=0)


def wl_ell_factor(ell):  # probes.py:73
    return np.sqrt((ell - 1) * (ell) * (ell + 1) * (ell + 2)) / (ell + 0.5) ** 2


def radial_kernels(bg, tracers, z):
    """The ell-independent part R_i(z) of every tracer kernel, [T, nz], and the
    per-tracer flag `is_wl` (WL tracers carry the probes.py:73 ell factor)."""
    c = bg.c
    a = 1.0 / (1.0 + z)
    Hz = H0 * np.sqrt(Esqr(c, a))
    # complex-step runs (oracle/derivatives.py): a perturbed gamma reaches R through the growth factor only, Omega_m / w0 / wa
    # through H and chi as well
    R = np.zeros((len(tracers), len(z)), dtype=np.result_type(Hz, bg.growth(a), bg.chi(a), np.float64))
    is_wl = np.zeros(len(tracers), dtype=bool)
    wl_idx = [i for i, t in enumerate(tracers) if t["kind"] == "wl"]
    # lensing efficiency is computed per probe in the reference; all WL tracers sharing the same
    # probe_zmax share the z' grid
    ext_idx = [i for i in wl_idx if tracers[i]["nz"]["family"] != "delta"]
    for pz in sorted(set(tracers[i]["probe_zmax"] for i in ext_idx)):
        idx = [i for i in ext_idx if tracers[i]["probe_zmax"] == pz]
        q = lensing_efficiency(bg, [tracers[i]["nz"] for i in idx], z, pz)
        for j, i in enumerate(idx):
            R[i] = q[j] * (3.0 * H0 ** 2 * c.Omega_m / 2.0 / C_LIGHT)  # probes.py:71
    chi = bg.chi(a)
    for i in wl_idx:  # delta_nz source planes (probes.py:53-64): no integral
        if tracers[i]["nz"]["family"] == "delta":
            chis = bg.chi(1.0 / (1.0 + np.array([tracers[i]["nz"]["params"][0]])))
            R[i] = _clip(chis - chi, 0, None) / _clip(chis, 1.0, None) * (1.0 + z) * chi * (
                3.0 * H0 ** 2 * c.Omega_m / 2.0 / C_LIGHT)
    for i, t in enumerate(tracers):
        if t["kind"] == "wl":
            is_wl[i] = True
            if t["ia"] is not None:  # probes.py:102-129
                R[i] += nz_eval(t["nz"], z) * bias_eval(t["ia"], bg, z) * Hz * (
                    -(C_1 * RHOCRIT) * c.Omega_m / bg.growth(a))
            R[i] *= 1.0 + t["m"]  # probes.py:205-207
        else:  # probes.py:77-99
            R[i] = nz_eval(t["nz"], z) * bias_eval(t["bias"], bg, z) * Hz
    return R, is_wl


# ----------------------------------------------------------------------------------------------
# angular_cl.py
# ----------------------------------------------------------------------------------------------
def cl_ordering(T):  # angular_cl.py:15-25
    return [(i, j) for i in range(T) for j in range(i, T)]


def pair_index(i, j, T):
    """Arithmetic form of _get_cov_blocks_ordering's find_index (angular_cl.py:34-38)."""
    i, j = (i, j) if i <= j else (j, i)
    return i * T - (i * (i - 1)) // 2 + (j - i)


def angular_cl(cosmo_row, ell, problem, stages=None):
    """angular_cl.py:49-98 -> [P, L].  `problem` = oracle.scenarios.flatten_spec(...)."""
    c = Cosmo(cosmo_row)
    c.transfer_type = problem.get("transfer", "eisenhu_osc")      # transfer.py:10 `type`
    c.prescription = problem.get("prescription", "takahashi2012")  # power.py:144 `prescription`
    bg = Background(c)
    pw = Power(bg)
    ell = np.atleast_1d(np.asarray(ell, dtype=np.float64))
    tracers = problem["tracers"]
    T = len(tracers)
    zmax = problem["zmax"]
    a = np.linspace(1.0 / (1.0 + zmax), 1.0, N_LIMBER + 1)
    wa = simps_weights(N_LIMBER) * ((1.0 - 1.0 / (1.0 + zmax)) / N_LIMBER) / 3
    z = 1.0 / a - 1.0
    chi = bg.chi(a)
    R, is_wl = radial_kernels(bg, tracers, z)
    geom = wa * dchioverda(c, a) / _clip(chi ** 2, 1.0, None) / C_LIGHT ** 2
    kk = (ell[:, None] + 0.5) / _clip(chi, 1.0, None)[None, :]  # [L, A]
    aa = np.broadcast_to(a[None, :], kk.shape)
    if problem["nonlinear"]:
        co = pw.halofit_coeffs(a)
        pklin = kk ** c.n_s * eisenstein_hu(c, kk) ** 2 * (bg.growth(a) ** 2)[None, :] * pw.pknorm
        d2l = kk ** 3 * pklin / (2.0 * np.pi ** 2)
        y = kk / co["k_nl"]
        d2q = d2l * ((1.0 + d2l) ** co["beta_n"] / (1 + co["alpha_n"] * d2l)) * np.exp(
            -(y / 4.0 + y ** 2 / 8.0))
        d2hp = co["a_n"] * y ** (3 * co["f1"]) / (
            1.0 + co["b_n"] * y ** co["f2"] + (co["c_n"] * co["f3"] * y) ** (3.0 - co["gamma_n"]))
        d2h = d2hp / (1.0 + co["mu_n"] / y + co["nu_n"] / y ** 2)
        pk = 2.0 * np.pi ** 2 / kk ** 3 * (d2q + d2h)
    else:
        pk = kk ** c.n_s * eisenstein_hu(c, kk) ** 2 * (bg.growth(a) ** 2)[None, :] * pw.pknorm
    del aa
    V = pk * geom[None, :]  # [L, A]
    ef = np.where(is_wl[:, None], wl_ell_factor(ell)[None, :], 1.0)  # [T, L]
    pairs = cl_ordering(T)
    ii = np.array([p[0] for p in pairs])
    jj = np.array([p[1] for p in pairs])
    KK = R[ii] * R[jj]  # [P, A]
    cl = (KK @ V.T) * ef[ii] * ef[jj]
    if stages is not None:
        stages.update(a=a, z=z, chi=chi, R=R, V=V, pk=pk, geom=geom, chitab=bg.chitab,
                      gtab=bg.gtab, pknorm=pw.pknorm, growth=bg.growth(a), sigmasqr8=sigmasqr_raw(c),
                      hubble=H0 * np.sqrt(Esqr(c, a)))
        if problem["nonlinear"]:
            stages.update(co)
    return cl


def noise_vector(problem):  # probes.py:210-223, 274-281
    out = []
    for t in problem["tracers"]:
        ng = t["nz"]["gals_per_arcmin2"] * STERADIAN_TO_ARCMIN2
        out.append(t["sigma_e"] ** 2 / ng if t["kind"] == "wl" else 1.0 / ng)
    return np.array(out)


def noise_cl(ell, problem):  # angular_cl.py:101-117
    T = len(problem["tracers"])
    nv = noise_vector(problem)
    L = len(np.atleast_1d(ell))
    return np.stack([nv[i] * (1.0 if i == j else 0.0) * np.ones(L) for i, j in cl_ordering(T)])


def gaussian_cl_covariance(ell, T, cl_signal, cl_noise, f_sky=0.25, sparse=True):
    """angular_cl.py:120-163."""
    ell = np.atleast_1d(np.asarray(ell, dtype=np.float64))
    L = len(ell)
    cl_obs = cl_signal + cl_noise
    P = cl_obs.shape[0]
    norm = (2 * ell + 1) * np.gradient(ell) * f_sky
    pairs = cl_ordering(T)
    cov = np.empty((P, P, L))
    for p, (i, j) in enumerate(pairs):
        for q, (m, n) in enumerate(pairs):
            cov[p, q] = (cl_obs[pair_index(i, m, T)] * cl_obs[pair_index(j, n, T)]
                         + cl_obs[pair_index(i, n, T)] * cl_obs[pair_index(j, m, T)]) / norm
    if sparse:
        return cov
    dense = cov[:, :, :, None] * np.eye(L)[None, None]
    return dense.transpose((0, 2, 1, 3)).reshape((L * P, L * P))


def gaussian_cl_covariance_and_mean(cosmo_row, ell, problem, f_sky=0.25, sparse=False):
    """angular_cl.py:166-196: returns (signal-only flattened mean, covariance)."""
    cl = angular_cl(cosmo_row, ell, problem)
    nl = noise_cl(ell, problem)
    cov = gaussian_cl_covariance(ell, len(problem["tracers"]), cl, nl, f_sky, sparse)
    return cl.flatten(), cov


# ----------------------------------------------------------------------------------------------
# likelihood.py / sparse.py (config 3 consumer)
# ----------------------------------------------------------------------------------------------
def sparse_to_dense(sp):
    """sparse.py:52-68: [ny, nx, n] block-diagonal-of-diagonals -> dense [ny*n, nx*n]."""
    ny, nx, n = sp.shape
    out = np.zeros((ny, n, nx, n))
    idx = np.arange(n)
    out[:, idx, :, idx] = np.moveaxis(sp, 2, 0)
    return out.reshape(ny * n, nx * n)


def gaussian_log_likelihood(data, mu, C, include_logdet=True):
    """likelihood.py:9-61 for a sparse covariance [P, P, L] (sparse.inv = per-ell inverse,
    sparse.py:295-315; slogdet = sum over ell, sparse.py:335-366), evaluated densely.  Keeps the
    reference's sign convention -0.5 * (chi2 - logdet) (likelihood.py:61)."""
    r = np.asarray(mu, dtype=np.float64) - np.asarray(data, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    P, _, L = C.shape
    if P * L <= 2000:  # literal dense evaluation
        dense = sparse_to_dense(C)
        chi2 = r @ np.linalg.solve(dense, r)
        logdet = np.linalg.slogdet(dense)[1]
    else:  # the dense matrix would be [P L, P L]: use its block structure (what sparse.inv / slogdet do)
        Cl = np.moveaxis(C, 2, 0)  # [L, P, P]
        rl = r.reshape(P, L).T  # [L, P]
        chi2 = float(np.sum(rl * np.linalg.solve(Cl, rl[:, :, None])[:, :, 0]))
        logdet = float(np.sum(np.linalg.slogdet(Cl)[1]))
    if not include_logdet:
        return -0.5 * chi2
    return -0.5 * (chi2 - logdet)
