"""Composite cone: host-side mirror of the reference cone operations the IP loop calls.

This is the *caller side* of the KKT path (the reference's Julia solver loop keeps these on
the host: `update_scaling!` runs at src/solver.jl:258-260 before `kkt_update!`).  The B200
backend reads the scaling *state* produced here (NN `w`; SOC `w, eta, d, u, v`; PSD `R`) and
derives the Hs blocks on the device; `get_Hs!` itself is therefore NOT implemented here
(it lives in the CUDA kernels and, for checking, in oracle/).

All cones of one type are processed together with numpy segment operations so that 1e4
second-order cones cost a handful of vector passes.

Reference files restated: src/cones/coneops_{zerocone,nncone,socone,psdtrianglecone}.jl,
coneops_compositecone.jl, compositecone_type.jl:96-141, coneops_symmetric_common.jl,
cone_api.jl:96-152 (collapsing), cone_types.jl:84-117 (SOC expansion threshold 4).
"""
import numpy as np

SOC_NO_EXPANSION_MAX_SIZE = 4          # cone_types.jl:101
ZERO, NONNEG, SOC, PSD, EXP, POW, GENPOW = 0, 1, 2, 3, 4, 5, 6
_NAMES = {"ZeroConeT": ZERO, "NonnegativeConeT": NONNEG,
          "SecondOrderConeT": SOC, "PSDTriangleConeT": PSD,
          "ExponentialConeT": EXP, "PowerConeT": POW, "GenPowerConeT": GENPOW}
_FLOATMAX = float(np.finfo(np.float64).max)
_SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


def ZeroConeT(dim): return ("ZeroConeT", int(dim))
def NonnegativeConeT(dim): return ("NonnegativeConeT", int(dim))
def SecondOrderConeT(dim): return ("SecondOrderConeT", int(dim))
def PSDTriangleConeT(dim): return ("PSDTriangleConeT", int(dim))
def ExponentialConeT(): return ("ExponentialConeT", 3)
def PowerConeT(alpha): return ("PowerConeT", 3, float(alpha))


def GenPowerConeT(alpha, dim2):
    alpha = tuple(float(a) for a in alpha)
    if abs(sum(alpha) - 1.0) > 1e-12 or min(alpha) <= 0:
        raise ValueError("GenPowerConeT: exponents must be positive and sum to 1")
    return ("GenPowerConeT", len(alpha) + int(dim2), alpha, int(dim2))


def triangular_number(k):
    return (k * (k + 1)) >> 1


def _nvars(spec):
    name, d = spec[0], spec[1]
    return triangular_number(d) if name == "PSDTriangleConeT" else d


def cones_new_collapsed(specs):
    """cone_api.jl:96-152: merge runs of NN / 1-dim SOC / 1-dim PSD, drop empty cones."""
    out = []
    it = list(specs)
    i = 0
    def collapsible(c):
        return (c[0] == "NonnegativeConeT" or
                (c[0] in ("SecondOrderConeT", "PSDTriangleConeT") and c[1] == 1))
    while i < len(it):
        c = it[i]; i += 1
        if _nvars(c) == 0:
            continue
        if collapsible(c):
            total = _nvars(c)
            while i < len(it):
                c2 = it[i]
                if _nvars(c2) == 0:
                    pass
                elif collapsible(c2):
                    total += _nvars(c2)
                else:
                    break
                i += 1
            out.append(NonnegativeConeT(total))
        else:
            out.append(c)
    return out


def _segsum(v, ptr):
    return np.add.reduceat(v, ptr[:-1]) if len(ptr) > 1 else np.zeros(0)


class CompositeCone:
    """compositecone_type.jl:8-65 + the per-type operations, vectorised by cone type."""

    def __init__(self, specs):
        self.specs = list(specs)
        nc = len(self.specs)
        self.types = np.array([_NAMES[s[0]] for s in self.specs], dtype=np.int32)
        self.dims = np.array([s[1] for s in self.specs], dtype=np.int64)      # SOC dim / PSD side n
        self.numels = np.array([_nvars(s) for s in self.specs], dtype=np.int64)
        for s in self.specs:
            if s[0] == "SecondOrderConeT" and s[1] < 2:
                raise ValueError("SOC dimension must be >= 2")
        self.rng_cones = np.concatenate([[0], np.cumsum(self.numels)]).astype(np.int64)
        self.numel = int(self.rng_cones[-1])
        # sparse-expandable SOCs (dim > 4) have a *diagonal* Hs block; so have generalised power
        # cones (always expanded, 3 extra columns: directldl_datamaps.jl:81-99)
        self.is_sparse = (self.types == SOC) & (self.dims > SOC_NO_EXPANSION_MAX_SIZE)
        self.is_genpow = self.types == GENPOW
        diag = (self.types == ZERO) | (self.types == NONNEG) | self.is_sparse | self.is_genpow
        blk = np.where(diag, self.numels, (self.numels * (self.numels + 1)) // 2)
        self.rng_blocks = np.concatenate([[0], np.cumsum(blk)]).astype(np.int64)
        self.Hs_is_diagonal = diag
        # nonsymmetric cones (one small object per cone: they are few and 3-dimensional)
        from . import nonsymmetric as _ns
        self.nonsym = []                                  # (cone index, object)
        for i, sp_ in enumerate(self.specs):
            if sp_[0] == "ExponentialConeT":
                self.nonsym.append((i, _ns.ExponentialCone()))
            elif sp_[0] == "PowerConeT":
                self.nonsym.append((i, _ns.PowerCone(sp_[2])))
            elif sp_[0] == "GenPowerConeT":
                self.nonsym.append((i, _ns.GenPowerCone(sp_[2], sp_[3])))
        deg = np.where(self.types == ZERO, 0,
              np.where(self.types == NONNEG, self.numels,
              np.where(self.types == SOC, 1,
              np.where(self.types == PSD, self.dims, 0))))
        self.degree = int(deg.sum()) + sum(c.degree for _, c in self.nonsym)
        # expansion columns appended to K, in cone order: 2 per sparse SOC, 3 per genpow cone
        self.pdims = np.where(self.is_sparse, 2, np.where(self.is_genpow, 3, 0)).astype(np.int64)
        self.p = int(self.pdims.sum())                   # pdim(sparse_maps)
        self.is_symmetric = len(self.nonsym) == 0
        self.allows_primal_dual_scaling = all(c.allows_primal_dual for _, c in self.nonsym)

        def _concat_ranges(sel):
            if not sel.any():
                return np.zeros(0, dtype=np.int64)
            return np.concatenate([np.arange(self.rng_cones[i], self.rng_cones[i + 1])
                                   for i in np.nonzero(sel)[0]])
        self.zero_idx = _concat_ranges(self.types == ZERO)
        self.nn_idx = _concat_ranges(self.types == NONNEG)
        # --- SOC space
        self.soc_cones = np.nonzero(self.types == SOC)[0]
        self.nsoc = len(self.soc_cones)
        sd = self.dims[self.soc_cones]
        self.soc_dims = sd
        self.soc_ptr = np.concatenate([[0], np.cumsum(sd)]).astype(np.int64)
        self.soc_idx = _concat_ranges(self.types == SOC)
        self.soc_head = self.soc_ptr[:-1]                      # positions inside soc space
        self.soc_tail = np.ones(len(self.soc_idx), dtype=bool)
        self.soc_tail[self.soc_head] = False
        self.soc_sparse = sd > SOC_NO_EXPANSION_MAX_SIZE
        # --- PSD groups by side length
        self.psd_groups = []
        for n in sorted(set(self.dims[self.types == PSD].tolist())):
            cn = np.nonzero((self.types == PSD) & (self.dims == n))[0]
            ne = triangular_number(n)
            idx = self.rng_cones[cn][:, None] + np.arange(ne)[None, :]
            ti, tj = np.tril_indices(n)
            rows, cols = tj, ti                                 # column-major packed upper
            diagpos = np.array([triangular_number(k + 1) - 1 for k in range(n)], dtype=np.int64)
            scale = np.where(rows == cols, 1.0, 1.0 / np.sqrt(2.0))
            self.psd_groups.append(dict(n=n, cones=cn, idx=idx, rows=rows, cols=cols,
                                        diagpos=diagpos, scale=scale))
        # --- scaling state (what the KKT backends read)
        m = self.numel
        self.w = np.zeros(m)            # NN: w ; SOC: w (unit hyperbolic vector)
        self.lam = np.zeros(m)          # NN/SOC: λ
        self.soc_eta = np.zeros(self.nsoc)
        self.soc_d = np.zeros(self.nsoc)
        self.soc_u = np.zeros(len(self.soc_idx))
        self.soc_v = np.zeros(len(self.soc_idx))
        for g in self.psd_groups:
            k, n = len(g["cones"]), g["n"]
            g["R"] = np.zeros((k, n, n)); g["Rinv"] = np.zeros((k, n, n))
            g["lam"] = np.zeros((k, n))

    # ------------------------------------------------------------------ helpers
    def _ns_items(self):
        """(cone object, slice into the stacked cone vectors) for every nonsymmetric cone."""
        for i, c in self.nonsym:
            yield c, slice(int(self.rng_cones[i]), int(self.rng_cones[i + 1]))

    def unit_initialization(self, z, s):
        """unit_initialization! (coneops_compositecone.jl:79-90): used instead of the symmetric
        default start whenever a nonsymmetric cone is present."""
        z[:] = 0.0; s[:] = 0.0
        z[self.nn_idx] = 1.0; s[self.nn_idx] = 1.0
        self.scaled_unit_shift(s, 1.0, "primal")          # SOC: e ; PSD: I   (NN handled above)
        self.scaled_unit_shift(z, 1.0, "dual")
        z[self.nn_idx] = 1.0; s[self.nn_idx] = 1.0
        for c, r in self._ns_items():
            z[r], s[r] = c.unit_initialization()

    def _soc_rep(self, percone):
        return np.repeat(percone, self.soc_dims)

    def _soc_residual(self, zs):
        z0 = zs[self.soc_head]
        n1 = np.sqrt(_segsum(np.where(self.soc_tail, zs * zs, 0.0), self.soc_ptr))
        return (z0 - n1) * (z0 + n1)

    def _soc_dot_tail(self, a, b):
        return _segsum(np.where(self.soc_tail, a * b, 0.0), self.soc_ptr)

    def _psd_mat(self, g, x):
        """svec_to_mat! (coneops_psdtrianglecone.jl:469-483), batched."""
        k, n = len(g["cones"]), g["n"]
        M = np.zeros((k, n, n))
        v = x[g["idx"]] * g["scale"][None, :]
        M[:, g["rows"], g["cols"]] = v
        M[:, g["cols"], g["rows"]] = v
        return M

    def _psd_svec(self, g, M):
        """mat_to_svec! (:486-497), batched; returns (k, numel)."""
        r, c = g["rows"], g["cols"]
        off = (M[:, r, c] + M[:, c, r]) / np.sqrt(2.0)
        return np.where((r == c)[None, :], M[:, r, c], off)

    # --------------------------------------------------------- composite ops
    def rectify_equilibration(self, delta, e):
        """coneops_compositecone.jl:29-47; NN/Zero elementwise (δ=1), others mean(e)/e."""
        delta[:] = 1.0
        changed = False
        for i in np.nonzero((self.types != ZERO) & (self.types != NONNEG))[0]:
            a, b = self.rng_cones[i], self.rng_cones[i + 1]
            delta[a:b] = e[a:b].mean() / e[a:b]
            changed = True
        return changed

    def margins(self, z, pd):
        alpha, beta = _FLOATMAX, 0.0
        if len(self.nn_idx):
            zn = z[self.nn_idx]
            alpha = min(alpha, float(zn.min()))
            beta += float(zn[zn > 0].sum())
        if self.nsoc:
            zs = z[self.soc_idx]
            a = zs[self.soc_head] - np.sqrt(self._soc_dot_tail(zs, zs))
            alpha = min(alpha, float(a.min()))
            beta += float(np.maximum(0.0, a).sum())
        for g in self.psd_groups:
            ev = np.linalg.eigvalsh(self._psd_mat(g, z))
            alpha = min(alpha, float(ev.min()))
            beta += float(ev[ev > 0].sum())
        return alpha, beta

    def scaled_unit_shift(self, z, alpha, pd):
        if pd == "primal":
            z[self.zero_idx] = 0.0
        z[self.nn_idx] += alpha
        if self.nsoc:
            z[self.soc_idx[self.soc_head]] += alpha
        for g in self.psd_groups:
            z[g["idx"][:, g["diagpos"]]] += alpha

    def set_identity_scaling(self):
        self.w[self.nn_idx] = 1.0
        if self.nsoc:
            ws = np.zeros(len(self.soc_idx)); ws[self.soc_head] = 1.0
            self.w[self.soc_idx] = ws
            self.soc_eta[:] = 1.0
            self.soc_d[:] = 0.5
            self.soc_u[:] = 0.0
            self.soc_u[self.soc_head] = np.sqrt(0.5)
            self.soc_v[:] = 0.0
        for g in self.psd_groups:
            g["R"][:] = np.eye(g["n"]); g["Rinv"][:] = np.eye(g["n"])

    def update_scaling(self, s, z, mu, strategy=0):
        """NT scaling.  NN: coneops_nncone.jl:77-89; SOC: coneops_socone.jl:75-154;
        PSD: coneops_psdtrianglecone.jl:78-143 (R, Rinv, λ; Hs is formed on the device).
        Nonsymmetric cones: dual or primal-dual scaling per `strategy` (nonsymmetric.py)."""
        for c, r in self._ns_items():
            try:
                if not c.update_scaling(s[r], z[r], mu, strategy):
                    return False
            except (ValueError, ZeroDivisionError, FloatingPointError):
                return False                              # z left the dual cone
        ni = self.nn_idx
        if len(ni):
            self.lam[ni] = np.sqrt(s[ni] * z[ni])
            self.w[ni] = np.sqrt(s[ni] / z[ni])
        if self.nsoc:
            zs, ss = z[self.soc_idx], s[self.soc_idx]
            rz, rs = self._soc_residual(zs), self._soc_residual(ss)
            if np.any(~(rz > 0)) or np.any(~(rs > 0)):
                return False
            zscale, sscale = np.sqrt(rz), np.sqrt(rs)
            self.soc_eta[:] = np.sqrt(sscale / zscale)
            w = ss / self._soc_rep(sscale)
            zr = zs / self._soc_rep(zscale)
            w = np.where(self.soc_tail, w - zr, w + zr)
            rw = self._soc_residual(w)
            if np.any(~(rw > 0)):
                return False
            wscale = np.sqrt(rw)
            w = w / self._soc_rep(wscale)
            w1sq = self._soc_dot_tail(w, w)
            w[self.soc_head] = np.sqrt(1.0 + w1sq)
            gam = 0.5 * wscale
            z0, s0 = zs[self.soc_head], ss[self.soc_head]
            c1 = (gam + z0 / zscale) / sscale
            c2 = (gam + s0 / sscale) / zscale
            lam = self._soc_rep(c1) * ss + self._soc_rep(c2) * zs
            lam = lam * self._soc_rep(1.0 / (s0 / sscale + z0 / zscale + 2 * gam))
            lam[self.soc_head] = gam
            lam = lam * self._soc_rep(np.sqrt(sscale * zscale))
            self.w[self.soc_idx] = w
            self.lam[self.soc_idx] = lam
            # sparse expansion data (computed for all; only read for dim > 4)
            w0 = w[self.soc_head]
            alpha = 2 * w0
            wsq = w0 * w0 + w1sq
            wsqinv = 1.0 / wsq
            d = wsqinv / 2
            u0 = np.sqrt(wsq - d)
            u1 = alpha / u0
            v1 = np.sqrt(2 * (2 + wsqinv) / (2 * wsq - wsqinv))
            self.soc_d[:] = d
            u = self._soc_rep(u1) * w; u[self.soc_head] = u0
            v = self._soc_rep(v1) * w; v[self.soc_head] = 0.0
            self.soc_u[:] = u; self.soc_v[:] = v
        for g in self.psd_groups:
            S, Z = self._psd_mat(g, s), self._psd_mat(g, z)
            try:
                L1 = np.linalg.cholesky(S); L2 = np.linalg.cholesky(Z)
            except np.linalg.LinAlgError:
                return False
            tmp = np.swapaxes(L2, 1, 2) @ L1
            U, sv, Vt = np.linalg.svd(tmp)
            g["lam"][:] = sv
            isq = 1.0 / np.sqrt(sv)
            g["R"][:] = (L1 @ np.swapaxes(Vt, 1, 2)) * isq[:, None, :]
            g["Rinv"][:] = isq[:, :, None] * (np.swapaxes(U, 1, 2) @ np.swapaxes(L2, 1, 2))
        return True

    def mul_Hs(self, y, x):
        """y = Hs x (coneops_compositecone.jl:138-150)."""
        y[self.zero_idx] = 0.0
        ni = self.nn_idx
        y[ni] = self.w[ni] * (self.w[ni] * x[ni])
        if self.nsoc:
            xs, w = x[self.soc_idx], self.w[self.soc_idx]
            c = 2 * _segsum(w * xs, self.soc_ptr)
            ys = xs.copy(); ys[self.soc_head] = -xs[self.soc_head]
            ys += self._soc_rep(c) * w
            ys *= self._soc_rep(self.soc_eta ** 2)
            y[self.soc_idx] = ys
        for g in self.psd_groups:
            tmp = self._psd_mul_W(g, x, g["R"], "N")
            y[g["idx"]] = self._psd_mul_W_vec(g, tmp, g["R"], "T")
        for c, r in self._ns_items():
            y[r] = c.mul_Hs(x[r])

    def _psd_mul_W(self, g, x, Rx, tr):
        return self._psd_mul_W_vec(g, x[g["idx"]], Rx, tr, gathered=True, xfull=x)

    def _psd_mul_W_vec(self, g, xv, Rx, tr, gathered=False, xfull=None):
        """mul_Wx_inner (coneops_psdtrianglecone.jl:409-437) on a (k,numel) svec batch."""
        k, n = len(g["cones"]), g["n"]
        X = np.zeros((k, n, n))
        v = xv * g["scale"][None, :]
        X[:, g["rows"], g["cols"]] = v
        X[:, g["cols"], g["rows"]] = v
        if tr == "T":
            Y = Rx @ X @ np.swapaxes(Rx, 1, 2)
        else:
            Y = np.swapaxes(Rx, 1, 2) @ X @ Rx
        return self._psd_svec(g, Y)

    def mul_W(self, tr, y, x):
        ni = self.nn_idx
        y[ni] = x[ni] * self.w[ni]
        if self.nsoc:
            y[self.soc_idx] = self._soc_mul_W(x[self.soc_idx], inv=False)
        for g in self.psd_groups:
            y[g["idx"]] = self._psd_mul_W_vec(g, x[g["idx"]], g["R"], tr)

    def mul_Winv(self, tr, y, x):
        ni = self.nn_idx
        y[ni] = x[ni] / self.w[ni]
        if self.nsoc:
            y[self.soc_idx] = self._soc_mul_W(x[self.soc_idx], inv=True)
        for g in self.psd_groups:
            y[g["idx"]] = self._psd_mul_W_vec(g, x[g["idx"]], g["Rinv"], tr)

    def _soc_mul_W(self, xs, inv):
        w = self.w[self.soc_idx]
        w0, x0 = w[self.soc_head], xs[self.soc_head]
        zeta = self._soc_dot_tail(w, xs)
        eta = self.soc_eta
        if not inv:
            c = x0 + zeta / (1 + w0)
            ys = self._soc_rep(eta) * (xs + self._soc_rep(c) * w)
            ys[self.soc_head] = eta * (w0 * x0 + zeta)
        else:
            c = -x0 + zeta / (1 + w0)
            ys = (xs + self._soc_rep(c) * w) / self._soc_rep(eta)
            ys[self.soc_head] = (w0 * x0 - zeta) / eta
        return ys

    def affine_ds(self, ds, s):
        ds[self.zero_idx] = 0.0
        ni = self.nn_idx
        ds[ni] = self.lam[ni] ** 2
        if self.nsoc:
            l = self.lam[self.soc_idx]
            ds[self.soc_idx] = self._soc_circ(l, l)
        for g in self.psd_groups:
            ds[g["idx"]] = 0.0
            ds[g["idx"][:, g["diagpos"]]] = g["lam"] ** 2
        for c, r in self._ns_items():
            ds[r] = c.affine_ds(s[r])

    def _soc_circ(self, y, z):
        x = self._soc_rep(y[self.soc_head]) * z + self._soc_rep(z[self.soc_head]) * y
        x[self.soc_head] = _segsum(y * z, self.soc_ptr)
        return x

    def combined_ds_shift(self, shift, step_z, step_s, sigma_mu):
        """_combined_ds_shift_symmetric! (coneops_symmetric_common.jl:2-36); overwrites
        step_z <- W step_z and step_s <- W^{-T} step_s like the reference."""
        shift[self.zero_idx] = 0.0
        # nonsymmetric cones read the unscaled affine step: sigma*mu*grad - eta (3rd-order term)
        for c, r in self._ns_items():
            shift[r] = c.combined_ds_shift(step_z[r].copy(), step_s[r].copy(), sigma_mu)
        tmp = step_z.copy()
        self.mul_W("N", step_z, tmp)
        tmp = step_s.copy()
        self.mul_Winv("T", step_s, tmp)
        ni = self.nn_idx
        shift[ni] = step_s[ni] * step_z[ni] - sigma_mu
        if self.nsoc:
            sh = self._soc_circ(step_s[self.soc_idx], step_z[self.soc_idx])
            sh[self.soc_head] -= sigma_mu
            shift[self.soc_idx] = sh
        for g in self.psd_groups:
            Y, Z = self._psd_mat(g, step_s), self._psd_mat(g, step_z)
            X = (Y @ Z + Z @ Y) / 2
            sv = self._psd_svec(g, X)
            sv[:, g["diagpos"]] -= sigma_mu
            shift[g["idx"]] = sv

    def ds_from_dz_offset(self, out, ds, work, z):
        """Δs_from_Δz_offset! (coneops_compositecone.jl:185-202)."""
        out[self.zero_idx] = 0.0
        ni = self.nn_idx
        out[ni] = ds[ni] / z[ni]
        if self.nsoc:
            zs, dss = z[self.soc_idx], ds[self.soc_idx]
            lam, w = self.lam[self.soc_idx], self.w[self.soc_idx]
            resz = self._soc_residual(zs)
            l1d1 = self._soc_dot_tail(lam, dss)
            w1d1 = self._soc_dot_tail(w, dss)
            o = -zs; o[self.soc_head] = zs[self.soc_head]
            c = lam[self.soc_head] * dss[self.soc_head] - l1d1
            o *= self._soc_rep(c / resz)
            eta, w0 = self.soc_eta, w[self.soc_head]
            head_val = o[self.soc_head] + eta * w1d1
            o += self._soc_rep(eta) * (dss + self._soc_rep(w1d1 / (1 + w0)) * w)
            o[self.soc_head] = head_val
            o *= self._soc_rep(1.0 / lam[self.soc_head])
            out[self.soc_idx] = o
        for g in self.psd_groups:
            # work = λ \ ds ; out = W^T work   (coneops_symmetric_common.jl:40-53)
            Zm = self._psd_mat(g, ds)
            l = g["lam"]
            X = 2 * Zm / (l[:, :, None] + l[:, None, :])
            wv = self._psd_svec(g, X)
            out[g["idx"]] = self._psd_mul_W_vec(g, wv, g["R"], "T")
        for c, r in self._ns_items():
            out[r] = ds[r]

    def compute_barrier(self, z, s, dz, ds, a):
        """compute_barrier (coneops_compositecone.jl:246-264) at (z + a dz, s + a ds)."""
        bar = 0.0
        ni = self.nn_idx
        if len(ni):
            v = (s[ni] + a * ds[ni]) * (z[ni] + a * dz[ni])
            if np.any(v <= 0):
                return np.inf
            bar -= float(np.log(v).sum())
        if self.nsoc:
            rs = self._soc_residual(s[self.soc_idx] + a * ds[self.soc_idx])
            rz = self._soc_residual(z[self.soc_idx] + a * dz[self.soc_idx])
            if np.any(rs <= 0) or np.any(rz <= 0):
                return np.inf
            bar -= float(np.log(rs * rz).sum()) / 2
        for g in self.psd_groups:
            for x, dx in ((z, dz), (s, ds)):
                try:
                    L = np.linalg.cholesky(self._psd_mat(g, x + a * dx))
                except np.linalg.LinAlgError:
                    return np.inf
                bar -= 2.0 * float(np.log(np.diagonal(L, axis1=1, axis2=2)).sum())
        for c, r in self._ns_items():
            bar += c.compute_barrier(z[r], s[r], dz[r], ds[r], a)
        return bar

    def step_length(self, dz, ds, z, s, alpha_max, settings=None):
        a = alpha_max
        ni = self.nn_idx
        if len(ni):
            for d, v in ((dz[ni], z[ni]), (ds[ni], s[ni])):
                neg = d < 0
                if neg.any():
                    a = min(a, float((-v[neg] / d[neg]).min()))
        if self.nsoc:
            a = min(a, self._soc_step(z[self.soc_idx], dz[self.soc_idx], a))
            a = min(a, self._soc_step(s[self.soc_idx], ds[self.soc_idx], a))
        for g in self.psd_groups:
            isq = 1.0 / np.sqrt(g["lam"])
            for tr, vec, Rx in (("N", dz, g["R"]), ("T", ds, g["Rinv"])):
                d = self._psd_mul_W_vec(g, vec[g["idx"]], Rx, tr)
                k, n = len(g["cones"]), g["n"]
                D = np.zeros((k, n, n))
                v = d * g["scale"][None, :]
                D[:, g["rows"], g["cols"]] = v; D[:, g["cols"], g["rows"]] = v
                D = D * isq[:, :, None] * isq[:, None, :]
                gam = float(np.linalg.eigvalsh(D).min())
                if gam < 0:
                    a = min(a, 1.0 / (-gam))
        if self.nonsym:
            # back off from full steps slightly so that centrality checks and logarithms do not
            # fail right at the boundary, then the nonsymmetric cones (compositecone.jl:230-243)
            a = min(a, 1.0 - _SQRT_EPS)
            back = settings.linesearch_backtrack_step if settings is not None else 0.8
            amin = settings.min_terminate_step_length if settings is not None else 1e-4
            for c, r in self._ns_items():
                az, as_ = c.step_length(dz[r], ds[r], z[r], s[r], a, amin, back)
                a = min(a, az, as_)
        return a, a

    def _soc_step(self, x, y, amax):
        """_step_length_soc_component (coneops_socone.jl:432-502), all cones at once."""
        x0, y0 = x[self.soc_head], y[self.soc_head]
        am = np.full(self.nsoc, amax)
        sel = (x0 >= 0) & (y0 < 0)
        am[sel] = np.minimum(am[sel], -x0[sel] / y0[sel])
        a = self._soc_residual(y)
        b = 2 * (x0 * y0 - self._soc_dot_tail(x, y))
        c = np.maximum(0.0, self._soc_residual(x))
        d = b * b - 4 * a * c
        out = am.copy()
        inf_step = ((a > 0) & (b > 0)) | (d < 0) | (a == 0)
        czero = (~inf_step) & (c == 0)
        out[czero] = np.where(a[czero] >= 0, am[czero], 0.0)
        gen = (~inf_step) & (c != 0)
        if gen.any():
            sq = np.sqrt(d[gen])
            t = np.where(b[gen] >= 0, -b[gen] - sq, -b[gen] + sq)
            r1 = (2 * c[gen]) / t
            r2 = t / (2 * a[gen])
            r1 = np.where(r1 < 0, _FLOATMAX, r1)
            r2 = np.where(r2 < 0, _FLOATMAX, r2)
            out[gen] = np.minimum(am[gen], np.minimum(r1, r2))
        return float(out.min())

    # ---------------------------------------------- state export for KKT backends
    def export_state(self):
        """Flat scaling state read by the KKT backends (H5 in SURVEY.md: upload the state,
        derive Hs on the device).  PSD R matrices are exported per cone in cone order,
        column-major n*n each."""
        Rlist = [None] * len(self.specs)
        for g in self.psd_groups:
            for j, ci in enumerate(g["cones"]):
                Rlist[ci] = g["R"][j]
        psdR = [np.asfortranarray(Rlist[i]).ravel(order="F")
                for i in np.nonzero(self.types == PSD)[0]]
        return dict(w=self.w, soc_eta=self.soc_eta, soc_d=self.soc_d,
                    soc_u=self.soc_u, soc_v=self.soc_v,
                    psd_R=(np.concatenate(psdR) if psdR else np.zeros(0)))

    def export_nonsymmetric_blocks(self):
        """K values of the nonsymmetric cones in the order the KKT maps expect them:
        hs   : per cone in cone order, packed triu of Hs (exp/pow, 6 values) or mu*(d1,d2) (genpow);
        q,r,p: the genpow expansion columns already scaled by -sqrt(mu)
               (_csc_update_sparsecone, directldl_datamaps.jl:146-166);  D = (-1,-1,+1) each."""
        hs, q, r, pp = [], [], [], []
        for i, c in self.nonsym:
            if self.types[i] == GENPOW:
                hs.append(c.hs_diag())
                sq = -np.sqrt(c.mu)
                q.append(sq * c.q); r.append(sq * c.r); pp.append(sq * c.p)
            else:
                hs.append(c.hs_triu())
        cat = lambda l: np.concatenate(l) if l else np.zeros(0)
        return dict(hs=cat(hs), q=cat(q), r=cat(r), p=cat(pp))
