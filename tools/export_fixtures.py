"""Write BASELINE workloads in the reference's JSON schema (src/json.jl) so that a Julia user can
time the true reference on exactly these instances:

    python tools/export_fixtures.py OUTDIR [C1 C2 C3 C4 C4r C5]
    julia> s = Clarabel.load_from_file("OUTDIR/C5.json"); Clarabel.solve!(s); s.timers

(C5 is ~0.6 GB of JSON; only the small C1 instance is committed, tests/golden/C1.json.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from clarabel_jl_b200 import problems  # noqa: E402

if __name__ == "__main__":
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    for w in sys.argv[2:] or ["C1"]:
        gen, kw, _, _ = bench.WORKLOADS[w]
        P, q, A, b, K = getattr(problems, gen)(**kw)
        path = os.path.join(out, f"{w}.json")
        problems.to_reference_json(path, P, q, A, b, K)
        print(path, os.path.getsize(path) >> 10, "KiB")
