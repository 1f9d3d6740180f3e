"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference KKT solver for parity checks
and the CPU baseline.  Never imported by the product.

Restates, statement for statement:
  * `DirectLDLKKTSolver` — src/kktsolvers/kktsolver_directldl.jl:5-94 (state), :211-245
    (`_kktsolver_update_inner!`), :247-310 (static regularisation + refactor), :313-371
    (setrhs / getlhs / solve), :389-466 (iterative refinement), :374-386 (update_P/A);
  * `_csc_update_sparsecone` for SOC — src/kktsolvers/direct-ldl/directldl_datamaps.jl:61-79;
  * `get_Hs!` — composite src/cones/coneops_compositecone.jl:123-132; NN coneops_nncone.jl:91-101;
    Zero coneops_zerocone.jl:91-102; SOC coneops_socone.jl:156-190; PSD
    coneops_psdtrianglecone.jl:153-161 with `skron!` :502-540 and `pack_triu` mathutils.jl:402;
  * the QDLDL engine wrapper — src/kktsolvers/direct-ldl/directldl_qdldl.jl (oracle/qdldl.py).

The KKT pattern / maps come from the product's host-side assembly (kkt_assembly.py), which is
itself checked against an independent scipy construction in tests/test_kkt_assembly.py.
"""
import importlib
import numpy as np

from . import qdldl as _q


def _pkg():
    return importlib.import_module("clarabel_jl_b200")


def skron_triu(A):
    """triu of the symmetric Kronecker product A (x)_s A in svec coordinates
    (coneops_psdtrianglecone.jl:502-540).  Vectorised: entry ((i,j),(k,l)) with i<=j, k<=l is
    c_ij c_kl (A_ik A_jl + A_il A_jk)/2 with c = 1 on the diagonal and sqrt(2) off it."""
    n = A.shape[0]
    ti, tj = np.tril_indices(n)
    I, J = tj, ti                       # column-major packed upper: i<=j
    c = np.where(I == J, 1.0, np.sqrt(2.0))
    out = (A[np.ix_(I, I)] * A[np.ix_(J, J)] + A[np.ix_(I, J)] * A[np.ix_(J, I)]) / 2.0
    out = out * c[:, None] * c[None, :]
    return np.triu(out)


def skron_triu_loops(A):
    """Literal loop form of skron! for small n (used to pin skron_triu)."""
    n = A.shape[0]
    ne = n * (n + 1) // 2
    out = np.zeros((ne, ne))
    s2 = np.sqrt(2.0)
    col = 0
    for l in range(n):
        for k in range(l + 1):
            row = 0
            kl = (k == l)
            for j in range(n):
                Ajl, Ajk = A[j, l], A[j, k]
                for i in range(j + 1):
                    if row > col:
                        break
                    ij = (i == j)
                    if not ij and not kl:
                        out[row, col] = A[i, k] * Ajl + A[i, l] * Ajk
                    elif ij and not kl:
                        out[row, col] = s2 * Ajl * Ajk
                    elif not ij and kl:
                        out[row, col] = s2 * A[i, l] * Ajk
                    else:
                        out[row, col] = Ajl * Ajl
                    row += 1
            col += 1
    return out


def get_Hs(cones, Hsblocks):
    """get_Hs!(cones, Hsblocks): one packed vector over cones.rng_blocks.  Diagonal blocks
    (Zero, NN, sparse SOC) are written with vector operations; dense SOC (dim <= 4) and PSD
    blocks cone by cone, literally as the reference does."""
    pk = _pkg().cones
    rc, rb = cones.rng_cones, cones.rng_blocks
    cache = getattr(cones, "_oracle_hs_cache", None)
    if cache is None:
        t = cones.types
        def blk_ranges(sel):
            ids = np.nonzero(sel)[0]
            if len(ids) == 0:
                return np.zeros(0, dtype=np.int64)
            return np.concatenate([np.arange(rb[i], rb[i + 1]) for i in ids])
        soc_order = {ci: k for k, ci in enumerate(cones.soc_cones)}
        sp_ids = np.nonzero(cones.is_sparse)[0]
        cache = dict(
            zero=blk_ranges(t == pk.ZERO), nn_blk=blk_ranges(t == pk.NONNEG),
            sp_blk=blk_ranges(cones.is_sparse),
            sp_head=rb[sp_ids].astype(np.int64),
            sp_k=np.array([soc_order[i] for i in sp_ids], dtype=np.int64),
            sp_dims=(rc[sp_ids + 1] - rc[sp_ids]).astype(np.int64),
            dense_ids=np.nonzero(((t == pk.SOC) & ~cones.is_sparse) | (t == pk.PSD))[0],
            soc_order=soc_order)
        cones._oracle_hs_cache = cache
    Hsblocks[cache["zero"]] = 0.0
    Hsblocks[cache["nn_blk"]] = cones.w[cones.nn_idx] ** 2
    if len(cache["sp_k"]):
        eta2 = cones.soc_eta[cache["sp_k"]] ** 2
        Hsblocks[cache["sp_blk"]] = np.repeat(eta2, cache["sp_dims"])
        Hsblocks[cache["sp_head"]] *= cones.soc_d[cache["sp_k"]]
    psd_seen = {}
    for i in cache["dense_ids"]:
        a, b = int(rc[i]), int(rc[i + 1])
        blk = Hsblocks[rb[i]:rb[i + 1]]
        if cones.types[i] == pk.SOC:
            soc_k = cache["soc_order"][i]
            eta2 = cones.soc_eta[soc_k] ** 2
            w = cones.w[a:b]
            dim = b - a
            blk[0] = (np.sqrt(2.0) * w[0] - 1.0) * (np.sqrt(2.0) * w[0] + 1.0)
            h = 1
            for col in range(1, dim):
                for row in range(col + 1):
                    blk[h] = 2 * w[row] * w[col]
                    h += 1
                blk[h - 1] += 1.0
            blk *= eta2
        else:
            n = int(cones.dims[i])
            g = next(g for g in cones.psd_groups if g["n"] == n)
            j = psd_seen.get(n, 0); psd_seen[n] = j + 1
            assert g["cones"][j] == i
            R = g["R"][j]
            Hs = skron_triu(R @ R.T)
            ti, tj = np.tril_indices(Hs.shape[0])
            blk[:] = Hs[tj, ti]                       # pack_triu: column-major upper
    # nonsymmetric cones: exp / pow store pack_triu(K.Hs) (coneops_expcone.jl:92-100,
    # coneops_powcone.jl), genpow the diagonal mu*(d1, d2) (coneops_genpowcone.jl:91-108)
    for i, c in getattr(cones, "nonsym", ()):
        blk = Hsblocks[rb[i]:rb[i + 1]]
        if cones.types[i] == pk.GENPOW:
            blk[:c.dim1] = c.mu * c.d1
            blk[c.dim1:] = c.mu * c.d2
        else:
            h = 0
            for col in range(3):
                for row in range(col + 1):
                    blk[h] = c.Hs[row, col]; h += 1
    return Hsblocks


class LinearSolverInfo:
    def __init__(self, name, threads, direct, nnzA, nnzL):
        self.name, self.threads, self.direct, self.nnzA, self.nnzL = name, threads, direct, nnzA, nnzL


class OracleDirectLDLKKTSolver:
    """DirectLDLKKTSolver{Float64} with the :qdldl engine, on the host."""

    literal_soc_updates = False     # True: issue the reference's 5 calls per sparse SOC
    amd_dense_scale = 1.5           # the reference's value (directldl_qdldl.jl:18-25); the full-size parity
                                    # tests lower it to shorten the host ordering (any permutation is valid)

    def __init__(self, P, A, cones, m, n, settings, perm=None):
        pkg = _pkg()
        self.m, self.n = m, n
        self.settings = settings
        self.KKT, self.map = pkg.kkt_assembly.assemble_kkt_matrix(P, A, cones)
        self.p = cones.p
        N = n + m + self.p
        self.x = np.zeros(N); self.b = np.zeros(N)
        self.work1 = np.zeros(N); self.work2 = np.zeros(N)
        self.Dsigns = pkg.kkt_assembly.fill_Dsigns(m, n, self.p, cones)
        self.Hsblocks = np.zeros(int(cones.rng_blocks[-1]))
        self.diagonal_regularizer = 0.0
        st = settings
        if perm is None and self.amd_dense_scale != 1.5:
            perm = _q.amd_order(self.KKT, self.amd_dense_scale)
        self.ldl = _q.QDLDLFactorisation(
            self.KKT, self.Dsigns, eps=st.dynamic_regularization_eps,
            delta=st.dynamic_regularization_delta, perm=perm,
            regularize=st.dynamic_regularization_enable)
        self._diag = self.KKT.diagonal  # noqa
        self.ir_rounds = 0; self.n_solves = 0; self.t_factor = 0.0; self.t_solve = 0.0

    # -- _update_values! / _scale_values!  (kktsolver_directldl.jl:130-188)
    def _update_values(self, index, values):
        self.KKT.data[index] = values
        self.ldl.update_values(index, values)

    def _scale_values(self, index, scale):
        self.KKT.data[index] *= scale
        self.ldl.scale_values(index, scale)

    def linear_solver_info(self):
        return LinearSolverInfo("qdldl", 1, True, self.ldl.nnzA, self.ldl.nnzL)

    def update(self, cones):
        mp = self.map
        get_Hs(cones, self.Hsblocks)
        self.Hsblocks *= -1.0
        self._update_values(mp.Hsblocks, self.Hsblocks)
        # sparse SOC expansions (directldl_datamaps.jl:61-79)
        if getattr(cones, "is_genpow", np.zeros(1, dtype=bool)).any():
            # generalised power cones (_csc_update_sparsecone, directldl_datamaps.jl:146-166):
            # columns q, r, p take the vectors, are scaled by -sqrt(mu), D = (-1, -1, +1)
            oq = orr = op = 0
            for k, (i, c) in enumerate((i, c) for i, c in cones.nonsym if cones.types[i] == _pkg().cones.GENPOW):
                iq = mp.gp_q[oq:oq + c.dim1]; ir = mp.gp_r[orr:orr + c.dim2]; ip = mp.gp_p[op:op + c.dim]
                self._update_values(iq, c.q); self._update_values(ir, c.r); self._update_values(ip, c.p)
                sq = -np.sqrt(c.mu)
                self._scale_values(iq, sq); self._scale_values(ir, sq); self._scale_values(ip, sq)
                self._update_values(mp.gp_D[3 * k:3 * k + 3], np.array([-1.0, -1.0, 1.0]))
                oq += c.dim1; orr += c.dim2; op += c.dim
        if cones.nsoc and cones.soc_sparse.any() and not self.literal_soc_updates:
            # same arithmetic as the per-cone loop below (nzval = u; nzval *= -eta^2 is bitwise
            # u * (-eta^2)), issued once over the concatenated maps so that 1e4 cones do not cost
            # 5e4 Python-level calls in the CPU baseline
            mask = np.repeat(cones.soc_sparse, cones.soc_dims)
            eta2 = cones.soc_eta[cones.soc_sparse] ** 2
            neg = -np.repeat(eta2, cones.soc_dims[cones.soc_sparse])
            self._update_values(mp.soc_u, cones.soc_u[mask] * neg)
            self._update_values(mp.soc_v, cones.soc_v[mask] * neg)
            D = np.empty(2 * len(eta2)); D[0::2] = -eta2; D[1::2] = eta2
            self._update_values(mp.soc_D, D)
        elif cones.nsoc and cones.soc_sparse.any():
            off = 0
            ks = 0
            sparse_ids = np.nonzero(cones.soc_sparse)[0]
            for k in sparse_ids:
                a, b = int(cones.soc_ptr[k]), int(cones.soc_ptr[k + 1])
                dim = b - a
                eta2 = cones.soc_eta[k] ** 2
                iu = mp.soc_u[off:off + dim]; iv = mp.soc_v[off:off + dim]
                self._update_values(iu, cones.soc_u[a:b])
                self._update_values(iv, cones.soc_v[a:b])
                self._scale_values(iu, -eta2)
                self._scale_values(iv, -eta2)
                self._update_values(mp.soc_D[2 * ks:2 * ks + 2], np.array([-eta2, eta2]))
                off += dim; ks += 1
        return self._regularize_and_refactor()

    def _regularize_and_refactor(self):
        import time
        st, mp = self.settings, self.map
        diag_kkt, diag_shifted = self.work1, self.work2
        if st.static_regularization_enable:
            diag_kkt[:] = self.KKT.data[mp.diag_full]
            maxdiag = float(np.abs(diag_kkt).max()) if len(diag_kkt) else 0.0
            eps = st.static_regularization_constant + st.static_regularization_proportional * maxdiag
            diag_shifted[:] = diag_kkt
            diag_shifted[self.Dsigns == 1] += eps
            diag_shifted[self.Dsigns != 1] -= eps
            self._update_values(mp.diag_full, diag_shifted)
            self.diagonal_regularizer = eps
        t = time.perf_counter()
        ok = self.ldl.refactor()
        self.t_factor += time.perf_counter() - t
        if st.static_regularization_enable:
            self.KKT.data[mp.diag_full] = diag_kkt
        return ok

    def setrhs(self, rhsx, rhsz):
        n, m = self.n, self.m
        self.b[:n] = rhsx
        self.b[n:n + m] = rhsz
        self.b[n + m:] = 0.0

    def _getlhs(self, lhsx, lhsz):
        n, m = self.n, self.m
        if lhsx is not None:
            lhsx[:] = self.x[:n]
        if lhsz is not None:
            lhsz[:] = self.x[n:n + m]

    def _ldl_solve(self, x, b):
        import time
        t = time.perf_counter()
        x[:] = b
        self.ldl.solve(x)
        self.t_solve += time.perf_counter() - t

    def _sym_mul(self, v):
        K = self.KKT
        return K @ v + K.T @ v - K.diagonal() * v

    def solve(self, lhsx, lhsz):
        self.n_solves += 1
        self._ldl_solve(self.x, self.b)
        if self.settings.iterative_refinement_enable:
            ok = self._iterative_refinement()
        else:
            ok = bool(np.all(np.isfinite(self.x)))
        if ok:
            self._getlhs(lhsx, lhsz)
        return ok

    def _refine_error(self, e, b, xi):
        e[:] = b - self._sym_mul(xi)
        return float(np.abs(e).max()) if len(e) else 0.0

    def _iterative_refinement(self):
        st = self.settings
        x, b = self.x, self.b
        e, dx = self.work1, self.work2
        normb = float(np.abs(b).max()) if len(b) else 0.0
        norme = self._refine_error(e, b, x)
        if not np.isfinite(norme):
            return False
        for _ in range(st.iterative_refinement_max_iter):
            if norme <= st.iterative_refinement_abstol + st.iterative_refinement_reltol * normb:
                break
            lastnorme = norme
            self._ldl_solve(dx, e)
            self.ir_rounds += 1
            dx += x
            norme = self._refine_error(e, b, dx)
            if not np.isfinite(norme):
                return False
            ratio = lastnorme / norme if norme > 0 else np.inf
            if ratio < st.iterative_refinement_stop_ratio:
                if ratio > 1.0:
                    x, dx = dx, x
                break
            x, dx = dx, x
        self.x, self.work2 = x, dx
        return True

    def update_P(self, P):
        self._update_values(self.map.P, P.data)

    def update_A(self, A):
        self._update_values(self.map.A, A.data)
