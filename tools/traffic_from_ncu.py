"""Turns an ncu launch list with dram__bytes_read.sum / dram__bytes_write.sum (see profiles/r02_summary.md
for the command) into the per-call DRAM traffic of the kernel classes bench.py reports:
    python tools/traffic_from_ncu.py gpurun_out/launches.csv C5 >> merges into profiles/r02_traffic.json"""
import csv, json, os, re, sys, collections
path, workload = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.defaultdict(lambda: collections.defaultdict(float))
for r in csv.DictReader(lines):
    nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("cb200::", "").replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    m = r["Metric Name"]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
        per[nm]["bytes"] += v
    elif m == "gpu__time_duration.sum":
        per[nm]["n"] += 1
def tot(pred):
    return sum(d["bytes"] for k, d in per.items() if pred(k))
nsolve = max(1, int(per["k_pack_perm"]["n"]))
nfac = max(1, int(per["k_scatter"]["n"]))
nres = max(1, int(sum(d["n"] for k, d in per.items() if k.startswith("k_residual<"))))
solve = lambda k: k.startswith(("k_fwd_", "k_bwd_", "k_big_", "k_pack_perm", "k_unpack_perm"))
out = dict(triangular_solve_sweeps=tot(solve) / nsolve,
           spmv_residual=tot(lambda k: k.startswith("k_residual")) / nres,
           schur_gemm=tot(lambda k: k.startswith("k_ldl_update")) / nfac,
           note=f"ncu dram__bytes_read.sum + dram__bytes_write.sum per call, {nsolve} solves / {nfac} factorisations / {nres} residuals in the capture window; "
                "schur_gemm = all k_ldl_update launches of one factorisation (panel updates + Schur complements)")
dst = os.path.join(ROOT, "profiles", "r02_traffic.json")
allw = json.load(open(dst)) if os.path.exists(dst) else {}
allw[workload] = out
json.dump(allw, open(dst, "w"), indent=1)
print(workload, {k: (round(v / 1e9, 3) if isinstance(v, float) else v) for k, v in out.items()})
