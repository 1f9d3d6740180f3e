"""Multi-GPU check (run under torchrun): the distributed factor/solve must reproduce the
single-GPU / oracle answers.  Usage: torchrun --nproc-per-node 2 tools/dist_check.py [C5s|C3s|...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import clarabel_jl_b200 as cb
from clarabel_jl_b200 import kktsolver_b200 as kb, lib
from common import small_instances, kkt_fixture, sym_full
from oracle import qdldl as oq
names = sys.argv[1:] or ["C5s", "C3s", "C2s", "C1"]
for name in names:
    KKT, mp, Ds, data, cones = kkt_fixture(cb, small_instances(cb)[name])
    N = KKT.shape[0]
    eng = kb.B200DirectLDLSolver(KKT, Ds, cb.Settings(), device=local, nd_leaf_size=32)
    if world > 1:
        kb.dist_init_from_torch(eng)
    ok = eng.refactor()
    F = oq.QDLDLFactorisation(KKT, Ds); F.refactor()
    rng = np.random.default_rng(3)
    err = 0.0
    for _ in range(3):
        b = rng.standard_normal(N)
        x = np.zeros(N); eng.solve(x, b)
        xo = b.copy(); F.solve(xo)
        err = max(err, np.abs(x - xo).max() / max(1.0, np.abs(xo).max()))
    S = lib.Symbolic(KKT, ordering=1, nd_leaf=32)
    owner, top, load = S.partition(world)
    print(f"[rank {rank}] {name}: N={N} refactor={ok} max rel err vs oracle={err:.2e} top fronts={int(top.sum())} "
          f"load={np.round(load / max(1, load.sum()), 3).tolist()}", flush=True)
    assert ok and err < 1e-9
if world > 1:
    dist.barrier(); dist.destroy_process_group()
print(f"[rank {rank}] dist_check ok")
