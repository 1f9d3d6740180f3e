"""Nonsymmetric cones (exponential, 3-d power, generalised power): host-side cone operations.

These are the caller side of the KKT path for the cone types whose `Hs` block is not an NT scaling
(reference: src/cones/coneops_expcone.jl, coneops_powcone.jl, coneops_genpowcone.jl,
coneops_nonsymmetric_common.jl).  The KKT backends only ever see the *result* of `update_scaling`:

  exp / pow : a dense symmetric 3x3 `Hs` (packed triu by get_Hs!, coneops_expcone.jl:92-100),
  genpow    : a diagonal `mu*(d1, d2)` plus three expansion columns `-sqrt(mu)*(q, r, p)` with
              D = (-1, -1, +1)  (coneops_genpowcone.jl:91-108, directldl_datamaps.jl:146-166).

Everything is written from the barrier functions

  exp dual : f*(z) = -log(z2 - z1 - z1 log(z3/-z1)) - log(-z1) - log(z3)
  pow dual : f*(z) = -log((z1/a)^(2a) (z2/(1-a))^(2-2a) - z3^2) - (1-a) log z1 - a log z2
  genpow   : f*(z) = -log(prod (z_i/a_i)^(2 a_i) - |w|^2) - sum (1-a_i) log z_i

with gradients, Hessians and the third-order correction derived from the generic chain rule for
F = -log(psi) (see `_third_order`); tests/test_nonsymmetric_cones.py checks all of them against
finite differences and the conjugacy identity  -grad f*(-grad f(s)) = s.
"""
import math
import numpy as np

_EPS = float(np.finfo(np.float64).eps)
_SQRT_EPS = math.sqrt(_EPS)
PRIMAL_DUAL, DUAL = 0, 1            # ScalingStrategy (types.jl:73-76)


def _logsafe(v):
    """mathutils.jl:12-18: log for v >= 0, -floatmax for negative arguments."""
    if v < 0:
        return -float(np.finfo(np.float64).max)
    if v == 0:
        return -math.inf
    return math.log(v)


def _safeguarded_root(h, dh, lo):
    """Root of a decreasing function h on (0, inf) with h(0+) = +inf and h(inf) < 0: Newton steps
    kept inside a sign bracket.  (The reference uses a one-sided Newton iteration from a
    closed-form starting point, coneops_nonsymmetric_common.jl:163-192; both converge to the same
    unique root.)"""
    while not h(lo) > 0:
        lo *= 0.5
        if lo < 1e-300:
            return lo
    hi = max(2 * lo, 1.0)
    while h(hi) > 0:
        lo, hi = hi, 2 * hi
        if hi > 1e300:
            return hi
    x = 0.5 * (lo + hi)
    for _ in range(200):
        fx = h(x)
        if fx > 0:
            lo = x
        else:
            hi = x
        d = dh(x)
        xn = x - fx / d if d != 0 else 0.5 * (lo + hi)
        if not (lo < xn < hi):
            xn = 0.5 * (lo + hi)
        if abs(xn - x) <= 4 * _EPS * abs(xn):
            return xn
        x = xn
    return x


def _third_order(psi, g, Hm, T_uv, sep3, u, v):
    """eta = +1/2 * D^3 F(z)[u, v] for F = -log(psi) + (separable logs) -- the sign the reference
    code computes (coneops_expcone.jl:296-341; its header comment writes -0.5, the arithmetic
    below it accumulates the bracket and divides by +2):
        D^3(-log psi)[u,v] = -2 g (g.u)(g.v)/psi^3 + (Hu (g.v) + Hv (g.u) + g (u'Hv))/psi^2 - T[u,v]/psi
    with g, Hm, T_uv the gradient, Hessian and contracted third derivative of psi; sep3 is the
    contracted third derivative of the separable part."""
    gu, gv = float(g @ u), float(g @ v)
    Hu, Hv = Hm @ u, Hm @ v
    d3 = (-2.0 * g * gu * gv / psi ** 3 + (Hu * gv + Hv * gu + g * float(u @ Hv)) / psi ** 2
          - T_uv / psi + sep3)
    return 0.5 * d3


class _Cone3:
    """Operations shared by the two 3-dimensional nonsymmetric cones."""
    dim = 3
    degree = 3
    numel = 3
    allows_primal_dual = True

    def __init__(self):
        self.grad = np.zeros(3)         # gradient of the dual barrier at the scaling point
        self.H_dual = np.zeros((3, 3))  # Hessian of the dual barrier there
        self.Hs = np.zeros((3, 3))
        self.z = np.zeros(3)

    # --- scaling (update_scaling!, coneops_expcone.jl:64-85 / coneops_powcone.jl)
    def update_scaling(self, s, z, mu, strategy):
        self.grad, self.H_dual = self.dual_grad_hess(z)
        if strategy == DUAL:
            self.Hs = mu * self.H_dual
        else:
            self._primal_dual_scaling(s, z)
        self.z = np.array(z, dtype=float)
        return True

    def _primal_dual_scaling(self, s, z):
        """Primal-dual (BFGS-type) scaling with Hs z = s and Hs z~ = s~, z~ = -f'(s), s~ = -f*'(z)
        (use_primal_dual_scaling, coneops_nonsymmetric_common.jl:77-160); falls back to mu H*
        close to the central path or when the secant conditions cannot be met."""
        H = self.H_dual
        st = self.grad
        zt = self.gradient_primal(s)
        dot_sz = float(s @ z)
        mu = dot_sz / 3.0
        mut = float(zt @ st) / 3.0
        ds = s + mu * st
        dz = z + mu * zt
        dot_dsz = float(ds @ dz)
        de1 = mu * mut - 1.0
        de2 = float(zt @ H @ zt) - 3.0 * mut * mut
        if abs(de1) > _SQRT_EPS and abs(de2) > _EPS and dot_sz > 0 and dot_dsz > 0:
            tmp = mut * st - H @ zt
            M = H - np.outer(st, st) / 3.0 - np.outer(tmp, tmp) / de2
            t = mu * float(np.linalg.norm(M))               # Frobenius norm
            axis = np.cross(z, zt)
            axis = axis / np.linalg.norm(axis)
            self.Hs = (np.outer(s, s) / dot_sz + np.outer(ds, ds) / dot_dsz + t * np.outer(axis, axis))
        else:
            self.Hs = mu * H

    # --- what the KKT backends read
    def hs_triu(self):
        """pack_triu(K.Hs): column-major upper triangle (6 values)."""
        Hs = self.Hs
        return np.array([Hs[0, 0], Hs[0, 1], Hs[1, 1], Hs[0, 2], Hs[1, 2], Hs[2, 2]])

    def mul_Hs(self, x):
        return self.Hs @ x

    def affine_ds(self, s):
        return np.array(s, dtype=float)

    def combined_ds_shift(self, step_z, step_s, sigma_mu):
        return self.grad * sigma_mu - self.higher_correction(step_s, step_z)

    def step_length(self, dz, ds, z, s, amax, amin, back):
        az = _backtrack(dz, z, amax, amin, back, self.is_dual_feasible)
        as_ = _backtrack(ds, s, amax, amin, back, self.is_primal_feasible)
        return az, as_

    def compute_barrier(self, z, s, dz, ds, a):
        return self.barrier_dual(z + a * dz) + self.barrier_primal(s + a * ds)

    def higher_correction(self, ds, v):
        """eta = 1/2 D^3 f*(z)[H*^-1 ds, v]; zero if H* is not numerically positive definite
        (coneops_expcone.jl:296-341)."""
        try:
            L = np.linalg.cholesky(self.H_dual)
        except np.linalg.LinAlgError:
            return np.zeros(3)
        u = np.linalg.solve(L.T, np.linalg.solve(L, ds))
        return self._eta(self.z, u, np.asarray(v, dtype=float))


def _backtrack(dq, q, a0, amin, back, inside):
    """backtrack_search (coneops_nonsymmetric_common.jl:5-34)."""
    a = a0
    while True:
        if inside(q + a * dq):
            return a
        a *= back
        if a < amin:
            return 0.0


class ExponentialCone(_Cone3):
    """K_exp = cl{ s : s3 >= s2 exp(s1/s2), s2 > 0 }."""

    def unit_initialization(self):
        s = np.array([-1.051383945322714, 0.556409619469370, 1.258967884768947])   # coneops_expcone.jl:45-47
        return s.copy(), s

    def is_primal_feasible(self, s):
        return bool(s[2] > 0 and s[1] > 0 and s[1] * _logsafe(s[2] / s[1]) - s[0] > 0)

    def is_dual_feasible(self, z):
        return bool(z[2] > 0 and z[0] < 0 and z[1] - z[0] - z[0] * _logsafe(-z[2] / z[0]) > 0)

    def barrier_dual(self, z):
        ell = _logsafe(-z[2] / z[0])
        return -_logsafe(-z[2] * z[0]) - _logsafe(z[1] - z[0] - z[0] * ell)

    def barrier_primal(self, s):
        w = _wright_omega(1.0 - s[0] / s[1] - _logsafe(s[1] / s[2]))
        w = (w - 1.0) * (w - 1.0) / w
        return -_logsafe(w) - 2.0 * _logsafe(s[1]) - _logsafe(s[2]) - 3.0

    def gradient_primal(self, s):
        w = _wright_omega(1.0 - s[0] / s[1] - _logsafe(s[1] / s[2]))
        g1 = 1.0 / ((w - 1.0) * s[1])
        g2 = g1 + g1 * _logsafe(w * s[1] / s[2]) - 1.0 / s[1]
        g3 = w / ((1.0 - w) * s[2])
        return np.array([g1, g2, g3])

    @staticmethod
    def _psi_parts(z):
        z1, z2, z3 = z
        ell = math.log(-z3 / z1)
        psi = z2 - z1 - z1 * ell
        g = np.array([-ell, 1.0, -z1 / z3])
        Hm = np.array([[1.0 / z1, 0.0, -1.0 / z3], [0.0, 0.0, 0.0], [-1.0 / z3, 0.0, z1 / (z3 * z3)]])
        return psi, g, Hm

    def dual_grad_hess(self, z):
        z1, z2, z3 = z
        psi, g, Hm = self._psi_parts(z)
        grad = -g / psi + np.array([-1.0 / z1, 0.0, -1.0 / z3])
        H = np.outer(g, g) / psi ** 2 - Hm / psi
        H[0, 0] += 1.0 / (z1 * z1)
        H[2, 2] += 1.0 / (z3 * z3)
        return grad, H

    def _eta(self, z, u, v):
        z1, z2, z3 = z
        psi, g, Hm = self._psi_parts(z)
        T = np.array([-u[0] * v[0] / z1 ** 2 + u[2] * v[2] / z3 ** 2, 0.0,
                      (u[0] * v[2] + u[2] * v[0]) / z3 ** 2 - 2.0 * z1 * u[2] * v[2] / z3 ** 3])
        sep3 = np.array([-2.0 * u[0] * v[0] / z1 ** 3, 0.0, -2.0 * u[2] * v[2] / z3 ** 3])
        return _third_order(psi, g, Hm, T, sep3, u, v)


def _wright_omega(beta):
    """Solution w of w + log(w) = beta for beta >= 1 (the exponential cone only needs that range)."""
    if beta < 0:
        raise ValueError(f"argument not in supported range: {beta}")
    w = beta - math.log(beta) if beta > 2.5 else 1.0 + 0.5 * (beta - 1.0)
    for _ in range(50):
        r = beta - w - math.log(w)
        # Halley step for phi(w) = w + log w - beta
        f1 = 1.0 + 1.0 / w
        f2 = -1.0 / (w * w)
        dw = r / (f1 + 0.5 * f2 * r / f1)
        w += dw
        if abs(dw) <= 2 * _EPS * abs(w):
            break
    return w


class PowerCone(_Cone3):
    """K_pow(a) = { s : s1^a s2^(1-a) >= |s3|, s1, s2 >= 0 }."""

    def __init__(self, alpha):
        super().__init__()
        self.alpha = float(alpha)

    def unit_initialization(self):
        a = self.alpha
        s = np.array([math.sqrt(1.0 + a), math.sqrt(1.0 + (1.0 - a)), 0.0])
        return s.copy(), s

    def is_primal_feasible(self, s):
        a = self.alpha
        return bool(s[0] > 0 and s[1] > 0 and
                    math.exp(2 * a * _logsafe(s[0]) + 2 * (1 - a) * _logsafe(s[1])) - s[2] * s[2] > 0)

    def is_dual_feasible(self, z):
        a = self.alpha
        return bool(z[0] > 0 and z[1] > 0 and
                    math.exp(2 * a * _logsafe(z[0] / a) + 2 * (1 - a) * _logsafe(z[1] / (1 - a))) - z[2] * z[2] > 0)

    def barrier_dual(self, z):
        a = self.alpha
        return (-_logsafe((z[0] / a) ** (2 * a) * (z[1] / (1 - a)) ** (2 - 2 * a) - z[2] * z[2])
                - (1 - a) * _logsafe(z[0]) - a * _logsafe(z[1]))

    def barrier_primal(self, s):
        # f(s) = -f*(-g(s)) - nu  with nu = 3
        g = self.gradient_primal(s)
        return -self.barrier_dual(-g) - 3.0

    def gradient_primal(self, s):
        """-z where z solves -grad f*(z) = s: eliminating z1, z2 leaves one equation in x = |z3|."""
        a = self.alpha
        s1, s2, s3 = s
        r = abs(s3)
        if r > _EPS:
            c1, c2 = (1 + a) / a, (2 - a) / (1 - a)
            ls = 2 * a * math.log(s1) + 2 * (1 - a) * math.log(s2)

            def h(x):
                return (2 * a * math.log(c1 + x * r) + 2 * (1 - a) * math.log(c2 + x * r) - ls
                        - math.log(x * x + 2 * x / r))

            def dh(x):
                return (2 * a * r / (c1 + x * r) + 2 * (1 - a) * r / (c2 + x * r)
                        - (2 * x + 2 / r) / (x * x + 2 * x / r))
            x = _safeguarded_root(h, dh, 1.0 / r)
            g3 = x if s3 > 0 else -x
            return np.array([-(a * g3 * s3 + 1 + a) / s1, -((1 - a) * g3 * s3 + 2 - a) / s2, g3])
        return np.array([-(1 + a) / s1, -(2 - a) / s2, 0.0])

    def _psi_parts(self, z):
        a = self.alpha
        z1, z2, z3 = z
        A, B = 2 * a, 2 - 2 * a
        phi = (z1 / a) ** A * (z2 / (1 - a)) ** B
        psi = phi - z3 * z3
        g = np.array([A * phi / z1, B * phi / z2, -2.0 * z3])
        Hm = np.array([[A * (A - 1) * phi / z1 ** 2, A * B * phi / (z1 * z2), 0.0],
                       [A * B * phi / (z1 * z2), B * (B - 1) * phi / z2 ** 2, 0.0],
                       [0.0, 0.0, -2.0]])
        return phi, psi, g, Hm

    def dual_grad_hess(self, z):
        a = self.alpha
        z1, z2, z3 = z
        phi, psi, g, Hm = self._psi_parts(z)
        grad = -g / psi + np.array([-(1 - a) / z1, -a / z2, 0.0])
        H = np.outer(g, g) / psi ** 2 - Hm / psi
        H[0, 0] += (1 - a) / (z1 * z1)
        H[1, 1] += a / (z2 * z2)
        return grad, H

    def _eta(self, z, u, v):
        a = self.alpha
        z1, z2, z3 = z
        A, B = 2 * a, 2 - 2 * a
        phi, psi, g, Hm = self._psi_parts(z)
        t111 = A * (A - 1) * (A - 2) * phi / z1 ** 3
        t112 = A * (A - 1) * B * phi / (z1 ** 2 * z2)
        t122 = A * B * (B - 1) * phi / (z1 * z2 ** 2)
        t222 = B * (B - 1) * (B - 2) * phi / z2 ** 3
        T = np.array([t111 * u[0] * v[0] + t112 * (u[0] * v[1] + u[1] * v[0]) + t122 * u[1] * v[1],
                      t112 * u[0] * v[0] + t122 * (u[0] * v[1] + u[1] * v[0]) + t222 * u[1] * v[1],
                      0.0])
        sep3 = np.array([-2.0 * (1 - a) * u[0] * v[0] / z1 ** 3, -2.0 * a * u[1] * v[1] / z2 ** 3, 0.0])
        return _third_order(psi, g, Hm, T, sep3, u, v)


class GenPowerCone:
    """K = { (u, w) : prod u_i^a_i >= |w|, u >= 0 }, u in R^dim1, w in R^dim2; dual scaling only,
    Hs = mu (diag(d1, d2 I) + p p' - q q' - r r') kept in factored form for the KKT expansion."""
    allows_primal_dual = False

    def __init__(self, alpha, dim2):
        self.alpha = np.asarray(alpha, dtype=float)
        self.dim1, self.dim2 = len(self.alpha), int(dim2)
        self.dim = self.numel = self.dim1 + self.dim2
        self.degree = self.dim1 + 1
        self.grad = np.zeros(self.dim)
        self.z = np.zeros(self.dim)
        self.mu = 1.0
        self.p = np.zeros(self.dim); self.q = np.zeros(self.dim1); self.r = np.zeros(self.dim2)
        self.d1 = np.zeros(self.dim1); self.d2 = 0.0

    def unit_initialization(self):
        s = np.concatenate([np.sqrt(1.0 + self.alpha), np.zeros(self.dim2)])
        return s.copy(), s

    def _phi(self, u, scaled):
        a = self.alpha
        return float(np.prod((u / a if scaled else u) ** (2 * a)))

    def is_primal_feasible(self, s):
        u, w = s[:self.dim1], s[self.dim1:]
        return bool(np.all(u > 0) and self._phi(u, False) - float(w @ w) > 0)

    def is_dual_feasible(self, z):
        u, w = z[:self.dim1], z[self.dim1:]
        return bool(np.all(u > 0) and self._phi(u, True) - float(w @ w) > 0)

    def barrier_dual(self, z):
        u, w = z[:self.dim1], z[self.dim1:]
        if not np.all(u > 0):
            return math.inf
        res = self._phi(u, True) - float(w @ w)
        return -_logsafe(res) - float(((1 - self.alpha) * np.log(u)).sum())

    def barrier_primal(self, s):
        return -self.barrier_dual(-self.gradient_primal(s)) - self.degree

    def gradient_primal(self, s):
        a = self.alpha
        u, w = s[:self.dim1], s[self.dim1:]
        nr = float(np.linalg.norm(w))
        g = np.empty(self.dim)
        if nr > _EPS:
            c = (1 + a) / a
            lu = float((2 * a * np.log(u)).sum())

            def h(x):
                return float((2 * a * np.log(x * nr + c)).sum()) - lu - math.log(2 * x / nr + x * x)

            def dh(x):
                return float((2 * a * nr / (nr * x + c)).sum()) - (2 * x + 2 / nr) / (x * x + 2 * x / nr)
            x = _safeguarded_root(h, dh, 1.0 / nr)
            g[self.dim1:] = x * w / nr
            g[:self.dim1] = -(1 + a + a * x * nr) / u
        else:
            g[self.dim1:] = 0.0
            g[:self.dim1] = -(1 + a) / u
        return g

    def update_scaling(self, s, z, mu, strategy):
        """update_dual_grad_H + mu (coneops_genpowcone.jl:66-84, 321-372)."""
        a = self.alpha
        u, w = z[:self.dim1], z[self.dim1:]
        phi = self._phi(u, True)
        n2 = float(w @ w)
        zeta = phi - n2
        if not zeta > 0:
            return False
        tau = 2 * a / u
        self.grad = np.concatenate([-tau * phi / zeta - (1 - a) / u, 2 * w / zeta])
        p0 = math.sqrt(phi * (phi + n2) / 2)
        p1 = -2 * phi / p0
        q0 = math.sqrt(zeta * phi / 2)
        r1 = 2 * math.sqrt(zeta / (phi + n2))
        self.d1 = tau * phi / (zeta * u) + (1 - a) / (u * u)
        self.d2 = 2 / zeta
        self.p = np.concatenate([p0 * tau / zeta, p1 * w / zeta])
        self.q = q0 * tau / zeta
        self.r = r1 * w / zeta
        self.mu = float(mu)
        self.z = np.array(z, dtype=float)
        return True

    def hs_diag(self):
        return self.mu * np.concatenate([self.d1, np.full(self.dim2, self.d2)])

    def mul_Hs(self, x):
        x1, x2 = x[:self.dim1], x[self.dim1:]
        y = np.concatenate([self.d1 * x1 - float(self.q @ x1) * self.q,
                            self.d2 * x2 - float(self.r @ x2) * self.r])
        y += float(self.p @ x) * self.p
        return self.mu * y

    def affine_ds(self, s):
        return np.array(s, dtype=float)

    def combined_ds_shift(self, step_z, step_s, sigma_mu):
        return self.grad * sigma_mu          # no higher-order term (coneops_genpowcone.jl:150-168)

    def step_length(self, dz, ds, z, s, amax, amin, back):
        az = _backtrack(dz, z, amax, amin, back, self.is_dual_feasible)
        as_ = _backtrack(ds, s, amax, amin, back, self.is_primal_feasible)
        return az, as_

    def compute_barrier(self, z, s, dz, ds, a):
        return self.barrier_primal(s + a * ds) + self.barrier_dual(z + a * dz)
