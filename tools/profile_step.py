"""Dev tool for ncu: builds a workload, runs one IP iteration, then N replay steps (graph replay
off so that every kernel is an individual launch ncu can see: CB200_GRAPH=0)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CB200_GRAPH", "0")
import numpy as np
import bench
import clarabel_jl_b200 as cb
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
P, q, A, b, K = bench.make_problem(name)
solver = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method="b200"))
ks = solver.kktsystem.kktsolver
rec = bench.Recorder(ks)
solver.solve(max_iter=1)
rec.detach()
st = rec.steps[-1]
lx, lz = np.zeros(solver.data.n), np.zeros(solver.data.m)
for i in range(nsteps):
    ks.update(bench.FakeCones(st["state"]))
    for rx, rz in st["rhs"]:
        ks.setrhs(rx, rz); ks.solve(lx, lz)
print("done", ks.ldl.timers(), ks.ldl.stats())
