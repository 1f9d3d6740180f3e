"""Dev tool (GPU): is the full-size SDP solve reproducible?  Runs C4 several times in ONE process
(after a C2 solve, with the oracle module imported like the test fixture does) and prints status,
iteration count, the tail of the iteration log, and whether two factorisations of the same recorded
system give bitwise identical pivots."""
import os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import clarabel_jl_b200 as cb
if os.environ.get("C4_IMPORT_ORACLE", "1") == "1":
    from oracle.kktsolver_oracle import OracleDirectLDLKKTSolver
    cb.register_kktsolver("qdldl", OracleDirectLDLKKTSolver)
try:
    from threadpoolctl import threadpool_info
    print("threadpools:", [(d.get("internal_api"), d.get("num_threads")) for d in threadpool_info()])
except Exception as e:
    print("threadpoolctl:", e)

def run(name):
    t = time.time()
    P, q, A, b, K = bench.make_problem(name)
    solver = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method="b200"))
    ks = solver.kktsystem.kktsolver
    rec = bench.Recorder(ks)
    t1 = time.time()
    sol = solver.solve()
    rec.detach()
    print(f"{name}: {sol.status_name} it={sol.iterations} setup {t1 - t:.1f}s solve {time.time() - t1:.1f}s "
          f"obj {sol.obj_val!r} r_prim {sol.r_prim:.3e} r_dual {sol.r_dual:.3e}", flush=True)
    for row in solver.iter_log[-3:]:
        print("    it %d pcost %.15e dcost %.15e pres %.3e dres %.3e mu %.3e step %.4f" % row)
    h = hashlib.sha1(np.ascontiguousarray(solver.iter_log[min(5, len(solver.iter_log) - 1)][1:3]).tobytes()).hexdigest()[:12]
    print("    hash of (pcost, dcost) at iteration 5:", h)
    full = [s for s in rec.steps if len(s["rhs"]) == 3]
    N = ks.KKT.shape[0]
    ds = []
    for _ in range(2):
        ks.update(bench.FakeCones(full[-1]["state"]))
        ds.append(ks.ldl.download(1, N).copy())
    print("    refactorisation bitwise reproducible:", bool(np.array_equal(ds[0], ds[1])), flush=True)
    return solver

if os.environ.get("C4_WARM", "1") == "1":
    run("C2")
for i in range(int(os.environ.get("C4_REPEAT", "3"))):
    s = run("C4")
    del s
