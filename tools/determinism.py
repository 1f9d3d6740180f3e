"""Dev tool (GPU): bitwise run-to-run reproducibility of the factorisation and of the solves.

    python tools/determinism.py C4r [repeats]

A few IP iterations give a realistic cone state; that system is then refactored `repeats` times and
solved for one fixed right-hand side each time.  Reported: how many repeats differ bitwise from the
first in D (pivots), in the panel storage (L, inverted diagonal blocks) and in the solution, plus the
largest relative difference.  Knobs to localise a difference: CB200_NO_TMA=1, CB200_MULTISTREAM=0,
CB200_GRAPH=0, CB200_MERGED_ROWS=0, CB200_TMA_TILE=128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import clarabel_jl_b200 as cb

name = sys.argv[1] if len(sys.argv) > 1 else "C4r"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 6
P, q, A, b, K = bench.make_problem(name)
solver = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method="b200"))
ks = solver.kktsystem.kktsolver
rec = bench.Recorder(ks)
solver.solve(max_iter=int(os.environ.get("CB200_MAX_ITER", "6")))
rec.detach()
full = [s for s in rec.steps if len(s["rhs"]) == 3]
st = full[-1]
N = ks.KKT.shape[0]; n, m = ks.n, ks.m
st_panel = int(ks.ldl.stats()["panel_bytes"] // 8)
rx, rz = st["rhs"][0]
ref = None
nd = dict(D=0, L=0, x=0); worst = dict(D=0.0, L=0.0, x=0.0)
import time
nzref = None; nz_bad = 0; upd_fail = 0
for r in range(R):
    if os.environ.get("DET_SLEEP"):
        time.sleep(float(os.environ["DET_SLEEP"]))        # let the GPU idle like it does between IP iterations
    if not ks.update(bench.FakeCones(st["state"])):
        upd_fail += 1; print(f"   repeat {r}: update reported failure", flush=True); continue
    nzv = ks.device_nzval()
    if nzref is None: nzref = nzv
    elif not np.array_equal(nzv, nzref):
        nz_bad += 1; print(f"   repeat {r}: device K values differ in {int((nzv != nzref).sum())} entries", flush=True)
    D = ks.ldl.download(1, N)
    L = ks.ldl.download(2, st_panel) if os.environ.get("DET_SKIP_L") != "1" else np.zeros(1)
    gx, gz = np.zeros(n), np.zeros(m)
    ks.setrhs(rx, rz); ks.solve(gx, gz)
    x = ks.ldl.download(6, N)
    if ref is None:
        ref = dict(D=D, L=L, x=x); continue
    for k, v in (("D", D), ("L", L), ("x", x)):
        if not np.array_equal(v, ref[k]):
            nd[k] += 1
            den = np.maximum(np.abs(ref[k]), 1e-300)
            worst[k] = max(worst[k], float(np.nanmax(np.abs(v - ref[k]) / den)))
            if k in ("D", "L"):
                bad = np.nonzero(v != ref[k])[0]
                big = bad[np.abs(v[bad] - ref[k][bad]) > 1e-12 * np.maximum(np.abs(ref[k][bad]), 1e-30)]
                print(f"   repeat {r}: {k} differs in {len(bad)} of {len(v)} entries, first {bad[0]} last {bad[-1]}; "
                      f"{len(big)} beyond 1e-12 relative" + (f", first of those {big[0]}" if len(big) else ""), flush=True)
knobs = {k: os.environ[k] for k in os.environ if k.startswith("CB200_")}
print(f"{name} {knobs}: repeats differing from the first (of {R - 1}): {nd}  worst rel diff {worst}  K-value mismatches {nz_bad}  failed updates {upd_fail}", flush=True)
