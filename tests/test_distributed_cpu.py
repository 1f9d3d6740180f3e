"""CPU (gloo, world_size 2): the multi-GPU host logic — subtree partition, ownership masks,
which contributions enter the all-reduce of the replicated top fronts, the gather of the solve —
emulated in numpy on the product's symbolic structures with torch.distributed collectives in the
places the CUDA path issues NCCL all-reduces (api_cuda.cu: allreduce_fronts / tri_solve)."""
import os
import socket
import numpy as np
from mf_numpy import panel_view
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import small_instances, kkt_fixture, sym_full


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _allreduce(a):
    t = torch.from_numpy(np.ascontiguousarray(a))
    dist.all_reduce(t)
    return t.numpy()


def _dist_factor_solve(rank, world, sym, owner, top, nzval, Ds, b):
    s = sym
    nsuper = len(s["sn_first"]) - 1
    N = len(s["perm"])
    mine = lambda sn: top[sn] or owner[sn] == rank
    active = lambda c: (rank == 0) if top[c] else (owner[c] == rank)
    L = np.zeros(int(s["panel_off"][-1])); U = np.zeros(int(s["upd_off"][-1])); D = np.zeros(N)
    np.add.at(L, s["a_map"], nzval)
    dsp = np.asarray(Ds)[s["perm"]]
    geo = []
    for sn in range(nsuper):
        f, l = int(s["sn_first"][sn]), int(s["sn_first"][sn + 1]); ns = l - f
        nr = int(s["rows_ptr"][sn + 1] - s["rows_ptr"][sn]); geo.append((f, l, ns, nr, ns + nr))
        if rank != 0 and top[sn]:
            panel_view(L, s, sn, ns + nr, ns)[:, :] = 0.0   # orig entries: rank 0 only
    for sn in range(nsuper):
        if not mine(sn):
            continue
        f, l, ns, nr, nf = geo[sn]
        F = np.zeros((nf, nf))
        F[:, :ns] = panel_view(L, s, sn, nf, ns).T
        for c in s["children"][sn]:
            if not active(c):
                continue
            nrc = geo[c][3]
            rel = s["rel"][s["rows_ptr"][c]:s["rows_ptr"][c + 1]]
            F[np.ix_(rel, rel)] += np.tril(U[s["upd_off"][c]:s["upd_off"][c] + nrc * nrc].reshape(nrc, nrc).T)
        if top[sn]:
            F = _allreduce(F)                       # root-front assembly across ranks
        for k in range(ns):
            d = F[k, k]; D[f + k] = d
            col = F[k + 1:, k].copy()
            F[k + 1:, k] = col / d
            F[k + 1:, k + 1:] -= np.tril(np.outer(col, col / d))
        panel_view(L, s, sn, nf, ns)[:, :] = F[:, :ns].T
        if nr:
            U[s["upd_off"][sn]:s["upd_off"][sn] + nr * nr] = np.tril(F[ns:, ns:]).T.reshape(-1)
    # ---- solve
    y = np.asarray(b, dtype=float)[s["perm"]].copy()
    u = np.zeros(len(s["rows"]))
    if rank != 0:
        for sn in range(nsuper):
            if top[sn]:
                y[geo[sn][0]:geo[sn][1]] = 0.0
    panel = lambda sn: panel_view(L, s, sn, geo[sn][4], geo[sn][2]).T
    for sn in range(nsuper):
        if not mine(sn):
            continue
        f, l, ns, nr, nf = geo[sn]
        w = np.zeros(nf); w[:ns] = y[f:l]
        for c in s["children"][sn]:
            if active(c):
                a0, a1 = s["rows_ptr"][c], s["rows_ptr"][c + 1]
                np.add.at(w, s["rel"][a0:a1], u[a0:a1])
        if top[sn]:
            w = _allreduce(w)
        P = panel(sn)
        for k in range(ns):
            w[k + 1:] -= P[k + 1:, k] * w[k]
        y[f:l] = w[:ns]; u[s["rows_ptr"][sn]:s["rows_ptr"][sn + 1]] = w[ns:]
    for sn in range(nsuper - 1, -1, -1):
        if not mine(sn):
            continue
        f, l, ns, nr, nf = geo[sn]
        rows = s["rows"][s["rows_ptr"][sn]:s["rows_ptr"][sn + 1]]
        P = panel(sn)
        w = np.concatenate([y[f:l] / D[f:l], y[rows]])
        for k in range(ns - 1, -1, -1):
            w[k] -= P[k + 1:, k] @ w[k + 1:]
        y[f:l] = w[:ns]
    for sn in range(nsuper):
        keep = (rank == 0) if top[sn] else (owner[sn] == rank)
        if not keep:
            y[geo[sn][0]:geo[sn][1]] = 0.0
    y = _allreduce(y)                               # gather of the per-rank solutions
    x = np.empty(N); x[s["perm"]] = y
    return x


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import clarabel_jl_b200 as cb
        from clarabel_jl_b200 import lib
        KKT, mp_, Ds, data, cones = kkt_fixture(cb, small_instances(cb)[name])
        S = lib.Symbolic(KKT, ordering=1, nd_leaf=32)
        a = S.arrays()
        owner, top, load = S.partition(world)
        rng = np.random.default_rng(11)
        b = rng.standard_normal(KKT.shape[0])
        x = _dist_factor_solve(rank, world, a, owner, top, KKT.data, Ds, b)
        res = np.abs(sym_full(KKT) @ x - b).max()
        out[rank] = (float(res), int(top.sum()), [float(v) for v in load])
    finally:
        dist.destroy_process_group()


def test_partition_covers_tree_and_balances(cb):
    from clarabel_jl_b200 import lib
    KKT, mp_, Ds, data, cones = kkt_fixture(cb, small_instances(cb)["C5s"])
    S = lib.Symbolic(KKT, ordering=1, nd_leaf=32)
    a = S.arrays()
    for world in (1, 2, 4, 8):
        owner, top, load = S.partition(world)
        par = a["sn_parent"]
        assert np.all(owner[top] == -1) and np.all((owner[~top] >= 0) & (owner[~top] < world))
        # top set is upward closed; subtrees are owned whole
        for sn in range(len(par)):
            if par[sn] >= 0:
                if top[sn]:
                    assert top[par[sn]]
                elif not top[par[sn]]:
                    assert owner[sn] == owner[par[sn]]
        if world > 1:
            assert top.sum() >= 1 and load.max() <= 0.75 * load.sum() + 1e-9


@pytest.mark.parametrize("name,world", [("C5s", 2), ("C3s", 2), ("C1s", 2), ("C3s", 4), ("C5s", 4), ("C5s", 8)])
def test_distributed_multifrontal_gloo(cb, name, world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert len(out) == world
    for r in range(world):
        res, ntop, load = out[r]
        assert res < 1e-9 and ntop >= 1
