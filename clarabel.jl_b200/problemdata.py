"""Problem data + Ruiz equilibration: mirror of src/problemdata.jl (setup-time, host).

Decides the *values* of P, A the KKT matrix sees, so it must match the reference
(src/problemdata.jl:3-88 constructor, :133-221 data_equilibrate!, src/utils/mathutils.jl
kkt_col_norms!/scale_data!).  The presolver (src/presolver.jl: rows of nonnegative cones whose
bound is infinite are removed, the solution is padded back with s = infinity, z = 0) is mirrored
because it changes the K the path sees; chordal decomposition is out of scope.
"""
import numpy as np
import scipy.sparse as sp

from .cones import cones_new_collapsed, CompositeCone

INFINITY = 1e20     # Clarabel.get_infinity() default


def _csc(M, shape=None):
    M = sp.csc_matrix(M, dtype=np.float64) if shape is None else sp.csc_matrix(M, shape=shape, dtype=np.float64)
    M.sort_indices()
    return M


class ProblemData:
    def __init__(self, P, q, A, b, cone_specs, settings):
        cones = cones_new_collapsed(cone_specs)
        P = _csc(P)
        # istriu check / triu copy (problemdata.jl:24-27); explicit zeros are kept
        coo = P.tocoo()
        keep = coo.row <= coo.col
        P = sp.csc_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=P.shape)
        P.sort_indices()
        A = _csc(A).copy()
        b = np.array(b, dtype=np.float64)
        # presolve (problemdata.jl:28-35, presolver.jl:107-147): only rows of NonnegativeCones
        self.presolve_keep = None
        self.mfull = len(b)
        if settings.presolve_enable:
            keep = np.ones(len(b), dtype=bool)
            bound = INFINITY * (1 - 10 * np.finfo(np.float64).eps)
            off = 0
            new_cones = []
            for c in cones:
                k = c[1] if c[0] != "PSDTriangleConeT" else (c[1] * (c[1] + 1)) // 2
                if c[0] == "NonnegativeConeT":
                    keep[off:off + k] = ~(b[off:off + k] > bound)
                    nkeep = int(keep[off:off + k].sum())
                    if nkeep > 0:
                        new_cones.append(("NonnegativeConeT", nkeep))
                else:
                    new_cones.append(c)
                off += k
            if not keep.all():
                self.presolve_keep = keep
                A = A[np.nonzero(keep)[0], :].tocsc(); A.sort_indices()
                b = b[keep]
                cones = new_cones
        self.P, self.A = P, A
        self.q = np.array(q, dtype=np.float64).copy()
        self.b = np.minimum(b, INFINITY)
        self.cones = cones
        self.m, self.n = A.shape
        self.dropped_zeros = False
        if settings.input_sparse_dropzeros:
            nz0 = self.P.nnz + self.A.nnz
            self.P.eliminate_zeros(); self.A.eliminate_zeros()
            self.dropped_zeros = (self.P.nnz + self.A.nnz) != nz0
        n, m = self.n, self.m
        self.d = np.ones(n); self.dinv = np.ones(n)
        self.e = np.ones(m); self.einv = np.ones(m)
        self.c = 1.0
        self.normq = float(np.abs(self.q).max()) if n else 0.0
        self.normb = float(np.abs(self.b).max()) if m else 0.0

    # ---- norms of unscaled data (problemdata.jl:91-111)
    def get_normq(self):
        if self.normq is None:                      # cleared by a data update: recover the unscaled norm
            self.normq = float(np.abs(self.q * self.dinv).max()) / self.c if self.n else 0.0
        return self.normq

    def get_normb(self):
        if self.normb is None:
            self.normb = float(np.abs(self.b * self.einv).max()) if self.m else 0.0
        return self.normb

    def equilibrate(self, cones: CompositeCone, settings):
        """data_equilibrate! (problemdata.jl:133-221)."""
        if not settings.equilibrate_enable:
            return
        P, A, q, b = self.P, self.A, self.q, self.b
        n, m = self.n, self.m
        smin, smax = settings.equilibrate_min_scaling, settings.equilibrate_max_scaling
        Pcol = np.repeat(np.arange(n), np.diff(P.indptr))
        Acol = np.repeat(np.arange(n), np.diff(A.indptr))
        d, e = self.d, self.e
        for _ in range(settings.equilibrate_max_iter):
            dwork = np.zeros(n); ework = np.zeros(m)
            # kkt_col_norms!: sym col norms of triu(P), col norms of A, row norms of A
            if P.nnz:
                ap = np.abs(P.data)
                np.maximum.at(dwork, Pcol, ap)
                np.maximum.at(dwork, P.indices, ap)
            if A.nnz:
                aa = np.abs(A.data)
                np.maximum.at(dwork, Acol, aa)
                np.maximum.at(ework, A.indices, aa)
            dwork[dwork == 0] = 1.0
            ework[ework == 0] = 1.0
            dwork = 1.0 / np.sqrt(dwork)
            ework = 1.0 / np.sqrt(ework)
            dwork = np.clip(dwork, smin / d, smax / d)
            ework = np.clip(ework, smin / e, smax / e)
            # scale_data!
            P.data *= dwork[P.indices] * dwork[Pcol]
            A.data *= ework[A.indices] * dwork[Acol]
            q *= dwork
            b *= ework
            d *= dwork
            e *= ework
            # cost scaling
            cn = np.zeros(n)
            if P.nnz:
                np.maximum.at(cn, Pcol, np.abs(P.data))
            mean_col_norm_P = cn.mean() if n else 0.0
            inf_norm_q = float(np.abs(q).max()) if n else 0.0
            if mean_col_norm_P != 0.0 and inf_norm_q != 0.0:
                scale_cost = max(inf_norm_q, mean_col_norm_P)
                ctmp = 1.0 / scale_cost
                ctmp = min(max(ctmp, smin / self.c), smax / self.c)
                P.data *= ctmp
                q *= ctmp
                self.c *= ctmp
        ework = np.ones(m)
        if cones.rectify_equilibration(ework, e):
            A.data *= ework[A.indices]
            b *= ework
            e *= ework
        self.dinv[:] = 1.0 / d
        self.einv[:] = 1.0 / e
