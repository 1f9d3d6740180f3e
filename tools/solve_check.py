"""Dev tool (GPU): whole IP solve of a workload on the GPU backend, then the recorded systems are
replayed and every solution is checked against the unregularised K by a host SpMV.
    python tools/solve_check.py C3 [nlast]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import bench
import clarabel_jl_b200 as cb
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
nlast = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P, q, A, b, K = bench.make_problem(name)
solver = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method="b200"))
ks = solver.kktsystem.kktsolver
rec = bench.Recorder(ks)
_mi = int(os.environ.get("CB200_MAX_ITER", "0"))
sol = solver.solve(max_iter=_mi) if _mi else solver.solve()
rec.detach()
for row in solver.iter_log[-5:]:
    print("   it %d pcost %.10e dcost %.10e pres %.2e dres %.2e mu %.2e step %.3f" % row)
print(name, sol.status_name, sol.iterations, "obj", sol.obj_val, "stats", {k: v for k, v in ks.ldl.stats().items() if k in ("ordering_used", "nlevels", "nsuper")})
n, m = ks.n, ks.m; N = ks.KKT.shape[0]
full = [(i, s) for i, s in enumerate(rec.steps) if len(s["rhs"]) == 3]
for i, st in full[-nlast:]:
    ok = ks.update(bench.FakeCones(st["state"]))
    Kd = sp.csc_matrix((ks.device_nzval(), ks.KKT.indices, ks.KKT.indptr), shape=ks.KKT.shape)
    mv = bench.sym_matvec(Kd)
    res = []
    for rx, rz in st["rhs"]:
        gx, gz = np.zeros(n), np.zeros(m)
        ks.setrhs(rx, rz); oks = ks.solve(gx, gz)
        xf = ks.ldl.download(6, N)
        bb = np.concatenate([rx, rz, np.zeros(N - n - m)])
        res.append((bool(oks), float(np.abs(bb - mv(xf)).max() / np.abs(bb).max())))
    nreg = int(ks.ldl.download(5, 1)[0])
    print(f"  iteration {i}: update ok={ok} nreg={nreg} residuals={res}")
    if nreg:
        idx = ks.ldl.download(7, 64).astype(np.int64); idx = idx[idx >= 0]
        Dv = ks.ldl.download(1, N)        # pivots in permuted order
        perm = ks.ldl.download(3, N).astype(np.int64)
        pos = np.empty(N, dtype=np.int64); pos[perm] = np.arange(N)
        diagK = Kd.diagonal()
        for j in idx:
            kind = "x" if j < n else ("z" if j < n + m else "expansion")
            extra = ""
            if kind == "z":
                ci = int(np.searchsorted(solver.cones.rng_cones, j - n, side="right") - 1)
                extra = f" cone {ci} type {int(solver.cones.types[ci])} dim {int(solver.cones.dims[ci])} offset {int(j - n - solver.cones.rng_cones[ci])}"
            print(f"     regularised pivot: original index {j} ({kind}{extra}) K_jj={diagK[j]:.3e} sign={int(ks.Dsigns[j])} D={Dv[pos[j]]:.3e}")
