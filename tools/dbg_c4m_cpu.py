import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import clarabel_jl_b200 as cb
from clarabel_jl_b200 import lib
from oracle.kktsolver_oracle import OracleDirectLDLKKTSolver
from oracle import qdldl as oq
cb.register_kktsolver("qdldl", OracleDirectLDLKKTSolver)
from common import small_instances
from mf_numpy import MFNumpy
P,q,A,b,K = small_instances(cb)["C4m"]()
s = cb.Solver(P,q,A,b,K,cb.Settings(direct_solve_method="qdldl")); s.solve(max_iter=10)
ks = s.kktsystem.kktsolver
# redo the last update manually to capture shifted matrix
s.cones.update_scaling(s.variables.s, s.variables.z, s.info.mu)
ks.update(s.cones)
Kd = ks.KKT.copy()
eps = ks.diagonal_regularizer
Kd.data[ks.map.diag_full] += np.where(ks.Dsigns==1, eps, -eps)
print("eps", eps, "qdldl nreg", ks.ldl.regularize_count, "diag range", np.abs(Kd.data[ks.map.diag_full]).min(), np.abs(Kd.data[ks.map.diag_full]).max())
for ordn in (0,1):
    S = lib.Symbolic(Kd, ordering=ordn, nd_leaf=96)
    mf = MFNumpy(S.arrays()); D = mf.factor(Kd.data, ks.Dsigns)
    print("ordering", ordn, "MFNumpy nreg", mf.nreg, "min|D|", np.abs(D).min(), S.stats["max_front"])
    F = oq.QDLDLFactorisation(Kd, ks.Dsigns, perm=S.arrays()["perm"]); F.refactor(); print("   qdldl w/ same perm nreg", F.regularize_count)
