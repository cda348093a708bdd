"""TEST INFRASTRUCTURE ONLY — CPU (numpy/scipy) restatement of AutoOED's GP-surrogate hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import this module.  The product package `autooed_b200` never does: it fails loudly when the
CUDA library is missing.

Parity status: PINNED for the GP / acquisition part — every function here is checked against the
reference's own modules executed in the authoring container (`oracle/make_golden.py` imports
`/root/reference/autooed/mobo/surrogate_model/gp.py` and `acquisition/*.py` unmodified and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them).  The reference ships no tests
or golden vectors of its own.

Each function cites the reference lines it restates (paths relative to /root/reference).
Third-party arithmetic on this path: scikit-learn `GaussianProcessRegressor`
(pinned 0.22.2.post1 in requirements.txt:11; `log_marginal_likelihood` = Rasmussen & Williams
Alg. 2.1 + eq. 5.9), SciPy `minimize(method='L-BFGS-B')`, `cdist`, `norm.pdf/cdf`.

Conventions: theta = log[c1, l_1..l_d, c2]; nu in {1, 3, 5, -1}
(Matern 1/2, 3/2, 5/2, RBF); k(r) = c1*phi(r) + c2; K = k(T,T) + 1e-10*I.
"""
import math

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.optimize import minimize
from scipy.spatial.distance import cdist
from scipy.special import ndtr

JITTER = 1e-10  # sklearn GaussianProcessRegressor(alpha=1e-10) default, used by gp.py:134
SQRT3 = math.sqrt(3.0)
SQRT5 = math.sqrt(5.0)
INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# model definition (gp.py:108-135)
# --------------------------------------------------------------------------------------
def initial_theta(d):
    """theta0 = log[1.0, 1..1, 1e-2]  (gp.py:126-132)."""
    return np.log(np.concatenate([[1.0], np.ones(d), [1e-2]]))


def theta_bounds(d):
    """log-bounds: c1, l in [sqrt(1e-3), sqrt(1e3)]; c2 in [e^-6, 1]  (gp.py:126-132)."""
    lo = np.concatenate([[math.log(math.sqrt(1e-3))], np.full(d, math.log(math.sqrt(1e-3))), [-6.0]])
    hi = np.concatenate([[math.log(math.sqrt(1e3))], np.full(d, math.log(math.sqrt(1e3))), [0.0]])
    return np.stack([lo, hi], axis=1)


def _split_theta(theta):
    theta = np.asarray(theta, dtype=np.float64)
    return math.exp(theta[0]), np.exp(theta[1:-1]), math.exp(theta[-1])


def _scaled_dist(A, B, ell):
    """r = cdist(A/l, B/l) by direct differences (gp.py:33,175) — not the |a|^2+|b|^2-2ab form."""
    return cdist(A / ell, B / ell, metric="euclidean")


def _phi(r, nu):
    """Kernel profile phi(r) (gp.py:36-43; RBF: sklearn kernels.RBF)."""
    if nu == 1:
        return np.exp(-r)
    if nu == 3:
        s = SQRT3 * r
        return (1.0 + s) * np.exp(-s)
    if nu == 5:
        s = SQRT5 * r
        return (1.0 + s + s * s / 3.0) * np.exp(-s)
    if nu == -1:
        return np.exp(-0.5 * r * r)
    raise ValueError(f"nu must be one of 1, 3, 5, -1 (got {nu})")


def _h(r, nu, phi=None):
    """h(r) with d phi / d log l_k = h(r) * D_k, D_k = ((x_k - t_k)/l_k)^2  (gp.py:63-79).
    nu=1: phi/r (0 where r == 0, gp.py:68-70); 3: 3 e^{-sqrt3 r}; 5: 5/3 (sqrt5 r + 1) e^{-sqrt5 r};
    RBF: e^{-r^2/2} (sklearn RBF: K * D_k)."""
    if nu == 1:
        phi = np.exp(-r) if phi is None else phi
        return np.divide(phi, r, out=np.zeros_like(r), where=(r != 0))
    if nu == 3:
        return 3.0 * np.exp(-SQRT3 * r)
    if nu == 5:
        s = SQRT5 * r
        return 5.0 / 3.0 * (s + 1.0) * np.exp(-s)
    return np.exp(-0.5 * r * r)


def _g_over_r(r, nu):
    """(d phi/d r) / r, the radial factor of d phi/d x_k = (g/r) * (x_k - t_k) / l_k^2
    (gp.py:181-191 with dd = (x-t)/(r l^2), gp.py:176-179).  nu=1 has a 1/r singularity:
    safe_divide returns 0 at r == 0 (gp.py:179)."""
    if nu == 1:
        return np.divide(-np.exp(-r), r, out=np.zeros_like(r), where=(r != 0))
    if nu == 3:
        return -3.0 * np.exp(-SQRT3 * r)
    if nu == 5:
        return -5.0 / 3.0 * np.exp(-SQRT5 * r) * (1.0 + SQRT5 * r)
    return -np.exp(-0.5 * r * r)


def kernel_matrix(theta, A, nu, B=None):
    """k(A, B) = c1*phi(r) + c2 (gp.py:24-58 composed with ConstantKernel product/sum, gp.py:130-132).
    For B=None the diagonal is exactly c1 + c2 (gp.py:55) and NO jitter is added here."""
    c1, ell, c2 = _split_theta(theta)
    r = _scaled_dist(A, A if B is None else B, ell)
    return c1 * _phi(r, nu) + c2


# --------------------------------------------------------------------------------------
# log marginal likelihood + gradient  (sklearn _gpr.py:541-655 called from gp.py:139)
# --------------------------------------------------------------------------------------
def lml_and_grad(theta, X, y, nu, want_grad=True):
    """Returns (lml, grad) with grad w.r.t. theta (log-parameters).
    Non-PD K -> (-inf, zeros) as sklearn does (_gpr.py:589-593)."""
    c1, ell, c2 = _split_theta(theta)
    n, d = X.shape
    r = _scaled_dist(X, X, ell)
    phi = _phi(r, nu)
    np.fill_diagonal(phi, 1.0)
    K = c1 * phi + c2
    K[np.diag_indices_from(K)] += JITTER
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return -np.inf, np.zeros_like(np.asarray(theta, dtype=np.float64))
    alpha = cho_solve((L, True), y, check_finite=False)
    lml = -0.5 * float(y @ alpha) - float(np.log(np.diag(L)).sum()) - 0.5 * n * math.log(2.0 * math.pi)
    if not want_grad:
        return lml, None
    Kinv = cho_solve((L, True), np.eye(n), check_finite=False)
    W = np.outer(alpha, alpha) - Kinv  # _gpr.py:627-633
    grad = np.empty(d + 2)
    grad[0] = 0.5 * float((W * (c1 * phi)).sum())
    G = W * (c1 * _h(r, nu, phi))
    for k in range(d):
        diff = (X[:, k, None] - X[None, :, k]) / ell[k]
        grad[1 + k] = 0.5 * float((G * diff * diff).sum())
    grad[d + 1] = 0.5 * c2 * float(W.sum())
    return lml, grad


def fit_theta(X, y, nu, theta0=None, bounds=None):
    """One L-BFGS-B run exactly as gp.py:94-101 (scipy defaults, result check disabled)."""
    d = X.shape[1]
    theta0 = initial_theta(d) if theta0 is None else np.asarray(theta0, dtype=np.float64)
    bounds = theta_bounds(d) if bounds is None else bounds

    def obj(th):
        lml, g = lml_and_grad(th, X, y, nu)
        return -lml, -g

    res = minimize(obj, theta0, method="L-BFGS-B", jac=True, bounds=bounds)
    return res.x, -res.fun, res.nfev


class GPState:
    """Fitted per-objective state (sklearn attributes kernel_.theta, L_, alpha_; _gpr.py:349-367)."""

    def __init__(self, theta, X, y, nu):
        self.theta = np.asarray(theta, dtype=np.float64)
        self.nu = nu
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        K = kernel_matrix(self.theta, self.X, nu)
        c1, ell, c2 = _split_theta(self.theta)
        np.fill_diagonal(K, c1 + c2)
        K[np.diag_indices_from(K)] += JITTER
        self.L = cholesky(K, lower=True, check_finite=False)  # raises LinAlgError if not PD (_gpr.py:352-361)
        self.alpha = cho_solve((self.L, True), self.y, check_finite=False)
        self._Linv = None
        self._Kinv = None

    @property
    def Linv(self):
        if self._Linv is None:
            self._Linv = solve_triangular(self.L, np.eye(self.L.shape[0]), lower=True, check_finite=False)
        return self._Linv

    @property
    def Kinv(self):
        """K^-1 = L^-T L^-1 formed explicitly, as gp.py:154-157 does."""
        if self._Kinv is None:
            Li_T = solve_triangular(self.L.T, np.eye(self.L.shape[0]), check_finite=False)
            self._Kinv = Li_T.dot(Li_T.T)
        return self._Kinv


def _safe_div(a, b):
    """utils/operand.py:8-12."""
    b = np.broadcast_to(b, np.broadcast(a, b).shape)
    return np.divide(a, b, out=np.zeros(b.shape), where=(b != 0))


# --------------------------------------------------------------------------------------
# posterior evaluate (gp.py:141-253), batched semantics = stack of per-row reference results
# --------------------------------------------------------------------------------------
def evaluate_normalized(states, Xs, std=False, gradient=False, hessian=False, mode="literal",
                        var_form="chol", chunk=256):
    """`GaussianProcess._evaluate` (gp.py:141-253) on normalised inputs; outputs in normalised-Y units.

    Batched semantics: every row is what the reference returns for that row alone (the reference's own
    batched std+gradient path is shape-broken for C > 1).

    mode='literal' reproduces the reference's Hessian broadcasting; mode='exact' gives
    the true second derivatives.  var_form='kinv' is the reference's explicit-inverse variance
    (gp.py:154-161); 'chol' is the triangular form k** - |L^-1 k*|^2 that the CUDA path executes.
    When hessian and std are requested the variance Hessian needs the gradient terms internally, so
    they are always computed (the reference raises UnboundLocalError for std+hessian without gradient).
    """
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    if Xs.ndim != 2:
        raise ValueError("X must be 2-D (rows, n_var)")
    C, d = Xs.shape
    m = len(states)
    F = np.empty((C, m))
    S = np.empty((C, m)) if std else None
    dF = np.empty((C, m, d)) if gradient else None
    dS = np.empty((C, m, d)) if (std and gradient) else None
    hF = np.empty((C, m, d, d)) if hessian else None
    hS = np.empty((C, m, d, d)) if (std and hessian) else None
    for o, st in enumerate(states):
        c1, ell, c2 = _split_theta(st.theta)
        T = st.X
        for s0 in range(0, C, chunk):
            x = Xs[s0:s0 + chunk]
            sl = slice(s0, s0 + x.shape[0])
            r = _scaled_dist(x, T, ell)
            k = c1 * _phi(r, st.nu) + c2  # gp.py:148
            F[sl, o] = k @ st.alpha  # gp.py:149
            if std:
                if var_form == "kinv":
                    v = k @ st.Kinv  # gp.py:160
                    var = (c1 + c2) - np.einsum("ij,ij->i", v, k)
                else:
                    w = k @ st.Linv.T
                    var = (c1 + c2) - np.einsum("ij,ij->i", w, w)
                    v = w @ st.Linv if (gradient or hessian) else None
                var = np.where(var < 0, 0.0, var)  # gp.py:163-165
                sd = np.sqrt(var)
                S[sl, o] = sd
            if not (gradient or hessian):
                continue
            gr = c1 * _g_over_r(r, st.nu)  # (c, N): dk/dx_k = gr * (x_k - t_k) / l_k^2
            if gradient or (std and hessian):
                # two-GEMM identity: sum_j w_j (x_k - t_jk) = (sum_j w_j) x_k - (w @ T)_k
                wa = gr * st.alpha
                dmu = (wa.sum(1, keepdims=True) * x - wa @ T) / ell ** 2  # gp.py:196
                if gradient:
                    dF[sl, o] = dmu
                if std:
                    wv = gr * v
                    dvar = -2.0 * (wv.sum(1, keepdims=True) * x - wv @ T) / ell ** 2  # gp.py:201-205
                    dsd = 0.5 * _safe_div(dvar, sd[:, None])  # gp.py:206 (per row)
                    if gradient:
                        dS[sl, o] = dsd
            if hessian:
                hmu, hvar = _hessian_rows(st, x, r, k, v if std else None, c1, ell, mode, std)
                hF[sl, o] = hmu
                if std:
                    if mode == "literal":  # gp.py:241: (hvar*sd - dvar_j*dsd_j) / var, product indexed by j only
                        num = hvar * sd[:, None, None] - (dvar * dsd)[:, None, :]
                    else:
                        num = hvar * sd[:, None, None] - dvar[:, :, None] * dsd[:, None, :]
                    hS[sl, o] = 0.5 * _safe_div(num, var[:, None, None])
    return {"F": F, "dF": dF, "hF": hF, "S": S, "dS": dS, "hS": hS}


def _hessian_rows(st, x, r, k, v, c1, ell, mode, std):
    """Second derivatives per row (gp.py:209-242).  Returns (hmu (c,d,d), hvar (c,d,d) or None)."""
    nu, T = st.nu, st.X
    c, d = x.shape
    N = T.shape[0]
    dx = x[:, None, :] - T[None, :, :]  # (c, N, d)
    rr = r[:, :, None]
    u = _safe_div(dx, rr * ell ** 2)  # dd, gp.py:176-179
    # hd_ij = (r delta_ij - dx_i u_j) / (r^2 l_i^2)   gp.py:212-214
    eye = np.eye(d)
    hd_num = rr[..., None] * eye[None, None] - dx[:, :, :, None] * u[:, :, None, :]
    hd_den = (rr ** 2)[..., None] * (ell ** 2)[None, None, :, None]
    hd = _safe_div(hd_num, hd_den)  # (c, N, d, d) [i, j]
    if mode == "literal":
        q = (u ** 2)[:, :, None, :]  # dd ** 2 broadcast: indexed by column j only
    else:
        q = u[:, :, :, None] * u[:, :, None, :]
    R = rr[..., None]
    if nu == 1:
        hK = -c1 * np.exp(-R) * (hd - q)
    elif nu == 3:
        hK = -3.0 * c1 * np.exp(-SQRT3 * R) * (R * hd + (1.0 - SQRT3 * R) * q)
    elif nu == 5:
        hK = -5.0 / 3.0 * c1 * np.exp(-SQRT5 * R) * (-5.0 * R ** 2 * q + (1.0 + SQRT5 * R) * (q + R * hd))
    else:
        hK = -c1 * np.exp(-0.5 * R ** 2) * ((1.0 - R ** 2) * q + R * hd)
    hmu = np.einsum("ctij,t->cij", hK, st.alpha)  # gp.py:230
    if not std:
        return hmu, None
    dK = (c1 * _g_over_r(r, nu))[:, :, None] * dx / ell ** 2  # (c, N, d)
    B = np.einsum("st,ctk->csk", st.Linv, dK)  # L^-1 dk_i
    if mode == "literal":  # 2 * dk_i K^-1 dk_i, index i for both factors (gp.py:236-240)
        mid = np.einsum("csi,csi->ci", B, B)[:, :, None]
    else:
        mid = np.einsum("csi,csj->cij", B, B)
    hvar = -(2.0 * np.einsum("ctij,ct->cij", hK, v) + 2.0 * mid)
    return hmu, hvar


# --------------------------------------------------------------------------------------
# normalisation (utils/normalization.py:24-35, 103-109; surrogate_model/base.py:31-52, 68-111)
# --------------------------------------------------------------------------------------
class Normalization:
    def __init__(self, xl, xu):
        self.xl, self.xu = np.asarray(xl, float), np.asarray(xu, float)
        self.y_mean = None
        self.y_scale = None

    def fit(self, Y):
        Y = np.asarray(Y, dtype=np.float64)
        self.y_mean = Y.mean(axis=0)
        sc = Y.std(axis=0)  # population std (sklearn StandardScaler)
        # sklearn _handle_zeros_in_scale: constant features (std below 10*eps*|mean|-ish) -> 1.0
        sc = np.where(sc < 10 * np.finfo(np.float64).eps, 1.0, sc)
        self.y_scale = sc

    def do_x(self, X):
        return np.clip((X - self.xl) / (self.xu - self.xl), 0, 1)

    def do_y(self, Y):
        return (Y - self.y_mean) / self.y_scale


def unnormalize(out, norm, std, gradient):
    """surrogate_model/base.py:106-109 — hF/hS are NOT rescaled."""
    out = dict(out)
    out["F"] = out["F"] * norm.y_scale + norm.y_mean
    if gradient:
        out["dF"] = out["dF"] * norm.y_scale[None, :, None]
    if std:
        out["S"] = out["S"] * norm.y_scale
    if std and gradient:
        out["dS"] = out["dS"] * norm.y_scale[None, :, None]
    return out


# --------------------------------------------------------------------------------------
# acquisitions (acquisition/{identity,ucb,ei,pi}.py)
# --------------------------------------------------------------------------------------
def _pdf(z):
    return INV_SQRT_2PI * np.exp(-0.5 * z * z)


def acquisition(kind, val, gradient=False, hessian=False, n_sample=None, y_min=None, mode="literal"):
    """kind in {'identity','ucb','ei','pi'}; val = un-normalised surrogate output dict.
    Returns (F, dF, hF) exactly as `<Acq>._evaluate` does."""
    mu, sd = val["F"], val["S"]
    dmu, hmu, dsd, hsd = val["dF"], val["hF"], val["dS"], val["hS"]
    if kind == "identity":  # identity.py:15-18
        return mu, dmu, hmu
    ex = lambda a: a[..., None]
    if kind == "ucb":  # ucb.py:19-53
        lam = math.sqrt(math.log(n_sample) / n_sample)
        F = mu - lam * sd
        dF = dmu - lam * dsd if gradient else None
        hF = hmu - lam * hsd if hessian else None
        return F, dF, hF
    y_min = np.asarray(y_min, dtype=np.float64)
    diff = y_min - mu
    z = _safe_div(diff, sd)
    pdf, cdf = _pdf(z), ndtr(z)
    dF = hF = None
    if kind == "ei":  # ei.py:20-70
        F = -diff * cdf - sd * pdf
        if gradient or hessian:
            dz_m = -_safe_div(1.0, sd)
            dz_s = -_safe_div(diff, sd ** 2)
            dpdf = -z * pdf
            if mode == "literal":
                a = cdf - diff * pdf * dz_m - sd * dpdf * dz_m  # ei.py:40
                b = diff * pdf * dz_s + pdf + sd * dpdf * dz_s  # ei.py:41 (sign quirk)
            else:
                a, b = cdf, -pdf
        if gradient:
            dF = ex(a) * dmu + ex(b) * dsd
        if hessian:
            if mode == "literal":  # ei.py:49-68
                dpdf_m, dpdf_s = dpdf * dz_m, dpdf * dz_s
                ddpdf_m = -z * dpdf_m - dz_m * pdf
                ddpdf_s = -z * dpdf_s - dz_s * pdf
                ddz_ss = _safe_div(diff, sd ** 3)
                h_m = -pdf * dz_m - dz_m * pdf + diff * dpdf * dz_m ** 2 + sd * dz_m * ddpdf_m
                h_s = diff * (dz_s * dpdf_s + pdf * ddz_ss) + dpdf_s + dpdf * dz_s + sd * dz_s * ddpdf_s + sd * dpdf * ddz_ss
                hF = (ex(ex(a)) * hmu + ex(ex(b)) * hsd
                      + ex(ex(h_m)) * dmu[..., :, None] * dmu[..., None, :]
                      + ex(ex(h_s)) * dsd[..., :, None] * dsd[..., None, :])
            else:
                inv = _safe_div(1.0, sd)
                f_mm, f_ms, f_ss = -pdf * inv, -pdf * z * inv, -z * z * pdf * inv
                hF = _exact_hess(a, b, f_mm, f_ms, f_ss, dmu, dsd, hmu, hsd)
        return F, dF, hF
    if kind == "pi":  # pi.py:20-64
        F = -cdf
        if gradient or hessian:
            dz_m = -_safe_div(1.0, sd)
            dz_s = -_safe_div(diff, sd ** 2)
            a = -pdf * dz_m
            b = -pdf * dz_s
        if gradient:
            dF = ex(a) * dmu + ex(b) * dsd
        if hessian:
            if mode == "literal":
                dpdf = -z * pdf
                hz_s = _safe_div(diff, sd ** 3)
                h_m = -(dpdf * dz_m) * dz_m
                h_s = -(dpdf * dz_s) * dz_s - pdf * hz_s
                hF = (ex(ex(a)) * hmu + ex(ex(b)) * hsd
                      + ex(ex(h_m)) * dmu[..., :, None] * dmu[..., None, :]
                      + ex(ex(h_s)) * dsd[..., :, None] * dsd[..., None, :])
            else:
                inv2 = _safe_div(1.0, sd ** 2)
                f_mm, f_ms, f_ss = z * pdf * inv2, (z * z - 1.0) * pdf * inv2, (z ** 3 - 2.0 * z) * pdf * inv2
                hF = _exact_hess(a, b, f_mm, f_ms, f_ss, dmu, dsd, hmu, hsd)
        return F, dF, hF
    raise ValueError(f"unknown acquisition {kind}")


def _exact_hess(a, b, f_mm, f_ms, f_ss, dmu, dsd, hmu, hsd):
    ex2 = lambda t: t[..., None, None]
    mm = dmu[..., :, None] * dmu[..., None, :]
    ss = dsd[..., :, None] * dsd[..., None, :]
    ms = dmu[..., :, None] * dsd[..., None, :] + dsd[..., :, None] * dmu[..., None, :]
    return ex2(a) * hmu + ex2(b) * hsd + ex2(f_mm) * mm + ex2(f_ms) * ms + ex2(f_ss) * ss


# --------------------------------------------------------------------------------------
# the whole surrogate as the reference's SurrogateModel.fit / evaluate pair
# --------------------------------------------------------------------------------------
class GPOracle:
    """CPU stand-in for `GaussianProcess` (gp.py:104-253) + `SurrogateModel` (base.py)."""

    def __init__(self, n_var, n_obj, xl=None, xu=None, nu=1):
        self.n_var, self.n_obj, self.nu = n_var, n_obj, nu
        self.norm = Normalization(np.zeros(n_var) if xl is None else xl, np.ones(n_var) if xu is None else xu)
        self.states = None
        self.nfev = []

    def fit(self, X, Y, thetas=None):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        self.norm.fit(Y)
        Xn, Yn = self.norm.do_x(X), self.norm.do_y(Y)
        self.states, self.nfev = [], []
        for o in range(self.n_obj):
            if thetas is None:
                th, _, nfev = fit_theta(Xn, Yn[:, o], self.nu)
                self.nfev.append(nfev)
            else:
                th = thetas[o]
            self.states.append(GPState(th, Xn, Yn[:, o], self.nu))
        return self

    def evaluate(self, X, std=False, gradient=False, hessian=False, mode="literal", var_form="chol",
                 normalized=False, chunk=256):
        X = np.asarray(X, dtype=np.float64)
        Xn = X if normalized else self.norm.do_x(X)
        out = evaluate_normalized(self.states, Xn, std, gradient, hessian, mode, var_form, chunk)
        out = unnormalize(out, self.norm, std, gradient)
        if mode == "exact" and hessian:  # exact mode also fixes B-3 (raw-Y units for hF, hS)
            out["hF"] = out["hF"] * self.norm.y_scale[None, :, None, None]
            if std:
                out["hS"] = out["hS"] * self.norm.y_scale[None, :, None, None]
        return out


# --------------------------------------------------------------------------------------
# extended-precision ground truth for the posterior variance
# --------------------------------------------------------------------------------------
def variance_truth_longdouble(theta, X, Xs, nu):
    """sigma^2 = k** - |L^-1 k*|^2 with kernel matrix, Cholesky factorisation and forward substitution carried out in
    numpy longdouble (80-bit on x86-64) — the yardstick for 'whose variance is closer to the truth' comparisons
    (gp.py:154-165 loses digits through the explicit inverse when cond(K) is large)."""
    ld = np.longdouble
    c1, ell, c2 = _split_theta(theta)
    c1, c2, ell = ld(c1), ld(c2), ell.astype(ld)
    A, B = X.astype(ld) / ell, Xs.astype(ld) / ell

    def kern(P, Q):
        r = np.sqrt(((P[:, None, :] - Q[None, :, :]) ** 2).sum(-1))
        s3, s5 = np.sqrt(ld(3)), np.sqrt(ld(5))
        if nu == 1:
            phi = np.exp(-r)
        elif nu == 3:
            phi = (1 + s3 * r) * np.exp(-s3 * r)
        elif nu == 5:
            phi = (1 + s5 * r + 5 * r * r / 3) * np.exp(-s5 * r)
        else:
            phi = np.exp(-r * r / 2)
        return c1 * phi + c2

    n = len(X)
    K = kern(A, A)
    K[np.diag_indices(n)] = c1 + c2 + ld(JITTER)
    L = np.zeros((n, n), dtype=ld)
    for j in range(n):
        L[j, j] = np.sqrt(K[j, j] - (L[j, :j] ** 2).sum())
        if j + 1 < n:
            L[j + 1:, j] = (K[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    ks = kern(B, A)  # (C, n)
    w = np.zeros_like(ks)
    for i in range(n):
        w[:, i] = (ks[:, i] - w[:, :i] @ L[i, :i]) / L[i, i]
    return ((c1 + c2) - (w ** 2).sum(axis=1)).astype(np.float64)
