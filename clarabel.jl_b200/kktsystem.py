"""Reduced 2x2 KKT driver: mirror of `DefaultKKTSystem` (src/kktsystem.jl).

Holds an AbstractKKTSolver (6-method interface, src/kktsolvers/kktsolver_defaults.jl:1-47):
    update(cones)->bool, setrhs(x,z), solve(lhsx|None, lhsz|None)->bool,
    update_P(P), update_A(A), linear_solver_info()
The concrete solver is chosen from settings.direct_solve_method through a small registry —
the one-line dispatch SURVEY.md section 8(b) says the reference needs at kktsystem.jl:33.
"""
import numpy as np

_KKT_SOLVER_REGISTRY = {}


def register_kktsolver(symbol, ctor):
    """ctor(P, A, cones, m, n, settings) -> AbstractKKTSolver-like object."""
    _KKT_SOLVER_REGISTRY[symbol] = ctor


def _quad_form_triu(x, P, y):
    """quad_form(x, Symmetric(P,:U), y)  (src/utils/mathutils.jl:299-337)."""
    if P.nnz == 0:
        return 0.0
    d = P.diagonal()
    return float(x @ (P @ y) + y @ (P @ x) - np.dot(d * x, y))


class DefaultKKTSystem:
    def __init__(self, data, cones, settings):
        m, n = data.m, data.n
        sym = settings.direct_solve_method
        if sym not in _KKT_SOLVER_REGISTRY:
            if sym == "b200":
                from . import kktsolver_b200  # noqa: F401  (registers itself; fails loudly w/o CUDA lib)
            else:
                raise ValueError(f"no KKT solver registered for direct_solve_method={sym!r}")
        self.kktsolver = _KKT_SOLVER_REGISTRY[sym](data.P, data.A, cones, m, n, settings)
        self.x1 = np.zeros(n); self.z1 = np.zeros(m)
        self.x2 = np.zeros(n); self.z2 = np.zeros(m)
        self.workx = np.zeros(n); self.workz = np.zeros(m)
        self.work_conic = np.zeros(m)

    def linear_solver_info(self):
        return self.kktsolver.linear_solver_info()

    def update(self, data, cones):
        """kkt_update! (kktsystem.jl:62-78)."""
        if not self.kktsolver.update(cones):
            return False
        return self._solve_constant_rhs(data)

    def _solve_constant_rhs(self, data):
        self.workx[:] = -data.q
        self.kktsolver.setrhs(self.workx, data.b)
        return self.kktsolver.solve(self.x2, self.z2)

    def solve_initial_point(self, variables, data):
        """kkt_solve_initial_point! (kktsystem.jl:95-132)."""
        ks = self.kktsolver
        if data.P.nnz == 0:
            self.workx[:] = 0.0
            self.workz[:] = data.b
            ks.setrhs(self.workx, self.workz)
            ok = ks.solve(variables.x, variables.s)
            variables.s *= -1.0
            if not ok:
                return ok
            self.workx[:] = -data.q
            self.workz[:] = 0.0
            ks.setrhs(self.workx, self.workz)
            ok = ks.solve(None, variables.z)
        else:
            self.workx[:] = -data.q
            self.workz[:] = data.b
            ks.setrhs(self.workx, self.workz)
            ok = ks.solve(variables.x, variables.z)
            variables.s[:] = -variables.z
        return ok

    def solve(self, lhs, rhs, data, variables, cones, steptype):
        """kkt_solve! (kktsystem.jl:135-215)."""
        x1, z1, x2, z2 = self.x1, self.z1, self.x2, self.z2
        workx, workz = self.workx, self.workz
        workx[:] = rhs.x
        ds_const = self.work_conic
        if steptype == "affine":
            ds_const[:] = variables.s
        else:
            cones.ds_from_dz_offset(ds_const, rhs.s, lhs.z, variables.z)
        workz[:] = ds_const - rhs.z
        self.kktsolver.setrhs(workx, workz)
        if not self.kktsolver.solve(x1, z1):
            return False
        xi = workx
        xi[:] = variables.x / variables.tau
        P = data.P
        tau_num = (rhs.tau - rhs.kappa / variables.tau + float(data.q @ x1) + float(data.b @ z1)
                   + 2 * _quad_form_triu(xi, P, x1))
        xi -= x2
        tau_den = variables.kappa / variables.tau - float(data.q @ x2) - float(data.b @ z2)
        tau_den += _quad_form_triu(xi, P, xi) - _quad_form_triu(x2, P, x2)
        lhs.tau = tau_num / tau_den
        lhs.x[:] = x1 + lhs.tau * x2
        lhs.z[:] = z1 + lhs.tau * z2
        cones.mul_Hs(lhs.s, lhs.z)
        lhs.s[:] = -(lhs.s + ds_const)
        lhs.kappa = -(rhs.kappa + variables.kappa * lhs.tau) / variables.tau
        return True

    def update_P(self, P):
        self.kktsolver.update_P(P)

    def update_A(self, A):
        self.kktsolver.update_A(A)
