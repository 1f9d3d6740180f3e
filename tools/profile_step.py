"""Dev tool for ncu: builds a workload, runs a couple of IP iterations, then N replay steps."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import clarabel_jl_b200 as cb
from clarabel_jl_b200 import problems
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
gen, kw, _, _ = bench.WORKLOADS[name]
P, q, A, b, K = getattr(problems, gen)(**kw)
solver = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method="b200"))
ks = solver.kktsystem.kktsolver
rec = bench.Recorder(ks, solver.cones)
solver.solve(max_iter=1)
rec.detach()
st = rec.steps[-1]
lx, lz = np.zeros(solver.data.n), np.zeros(solver.data.m)
import ctypes
cudart = ctypes.CDLL("libcudart.so") if False else None
for i in range(nsteps):
    ks.update(bench.FakeCones(st["state"]))
    for rx, rz in st["rhs"]:
        ks.setrhs(rx, rz); ks.solve(lx, lz)
print("done", ks.ldl.timers())
