"""Test helper: numpy emulation of the multifrontal numeric phase on the product's symbolic
structures (panels, update blocks, a_map, rel).  It exercises exactly the index maps the CUDA
kernels use, so symbolic-analysis bugs are caught on the CPU.  Not part of the product."""
import numpy as np


def panel_view(L, s, sn, nf, ns):
    """(ns, nf) view (transposed panel) of supernode sn; the panel is column-major with leading
    dimension panel_ld[sn] (>= nf: large fronts are padded)."""
    ld = int(s["panel_ld"][sn]) if "panel_ld" in s else nf
    off = int(s["panel_off"][sn])
    return L[off:off + ld * ns].reshape(ns, ld)[:, :nf]


class MFNumpy:
    def __init__(self, sym):
        """sym: dict of arrays from cb200_symbolic_get (see lib.Symbolic.arrays())."""
        self.s = sym
        self.N = len(sym["perm"])
        self.nsuper = len(sym["sn_first"]) - 1

    def factor(self, nzval, dsigns_orig, eps=1e-13, delta=2e-7):
        s = self.s
        L = np.zeros(int(s["panel_off"][-1]))
        U = np.zeros(int(s["upd_off"][-1]))
        np.add.at(L, s["a_map"], nzval)
        D = np.zeros(self.N)
        dsp = np.asarray(dsigns_orig)[s["perm"]]
        self.nreg = 0
        for sn in range(self.nsuper):
            f, l = int(s["sn_first"][sn]), int(s["sn_first"][sn + 1])
            ns = l - f
            nr = int(s["rows_ptr"][sn + 1] - s["rows_ptr"][sn])
            nf = ns + nr
            F = np.zeros((nf, nf))
            F[:, :ns] = panel_view(L, s, sn, nf, ns).T
            for c in s["children"][sn]:
                nrc = int(s["rows_ptr"][c + 1] - s["rows_ptr"][c])
                rel = s["rel"][s["rows_ptr"][c]:s["rows_ptr"][c + 1]]
                Uc = U[s["upd_off"][c]:s["upd_off"][c] + nrc * nrc].reshape(nrc, nrc).T
                F[np.ix_(rel, rel)] += np.tril(Uc)
            for k in range(ns):
                d = F[k, k]
                if d * dsp[f + k] < eps:
                    d = delta * dsp[f + k]; self.nreg += 1
                D[f + k] = d
                col = F[k + 1:, k].copy()
                F[k + 1:, k] = col / d
                F[k + 1:, k + 1:] -= np.tril(np.outer(col, col / d))
                F[k, k] = 1.0
            panel_view(L, s, sn, nf, ns)[:, :] = F[:, :ns].T
            if nr:
                U[s["upd_off"][sn]:s["upd_off"][sn] + nr * nr] = np.tril(F[ns:, ns:]).T.reshape(-1)
        self.L, self.D = L, D
        return D

    def solve(self, b):
        s = self.s
        y = np.asarray(b, dtype=float)[s["perm"]].copy()
        # forward (right-looking form; equivalent to the multifrontal form the kernels use)
        for sn in range(self.nsuper):
            f, l = int(s["sn_first"][sn]), int(s["sn_first"][sn + 1])
            ns = l - f
            rows = s["rows"][s["rows_ptr"][sn]:s["rows_ptr"][sn + 1]]
            nf = ns + len(rows)
            P = panel_view(self.L, s, sn, nf, ns).T
            for k in range(ns):
                y[f + k + 1:l] -= P[k + 1:ns, k] * y[f + k]
            y[rows] -= P[ns:, :] @ y[f:l]
        y /= self.D
        for sn in range(self.nsuper - 1, -1, -1):
            f, l = int(s["sn_first"][sn]), int(s["sn_first"][sn + 1])
            ns = l - f
            rows = s["rows"][s["rows_ptr"][sn]:s["rows_ptr"][sn + 1]]
            nf = ns + len(rows)
            P = panel_view(self.L, s, sn, nf, ns).T
            y[f:l] -= P[ns:, :].T @ y[rows]
            for k in range(ns - 1, -1, -1):
                y[f + k] -= P[k + 1:ns, k] @ y[f + k + 1:l]
        x = np.empty(self.N)
        x[s["perm"]] = y
        return x
