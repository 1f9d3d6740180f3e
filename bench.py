#!/usr/bin/env python
"""bench.py — IP-iterations/sec of the KKT path (BASELINE.json metric).

One "step" = the KKT work of one interior-point iteration, issued through the reference-facing
AbstractKKTSolver boundary exactly as `solve!` issues it (src/solver.jl:278-323):
    kktsolver_update!(cones)                    value scatter + static reg + numeric LDL'
    3 x (kktsolver_setrhs! ; kktsolver_solve!)  constant-RHS, affine and combined solves, each
                                                with iterative refinement
on cone states / right-hand sides recorded from a real IP run of the synthetic instance.

  value : steps/s with the recorded inputs already resident in HBM (cb200_set_resident)
  e2e   : steps/s through the C-ABI with HOST buffers (H2D of the cone state and the three
          right-hand sides and D2H of the three solutions inside the timed region)
Timing: CUDA events on the library's own stream, max over ranks, W warm-up steps first.  The
factor/solve working set (panel storage) is far larger than the 126 MB L2 for every workload
except C1, so no explicit L2 flush is done ("inputs larger than L2").

--impl reference : the reference's CPU path (QDLDL-algorithm restatement, 1 thread like
directldl_qdldl.jl:37) on a bounded sample of the same workload; the measured sample rate is
converted to iterations of the FULL workload per second with the per-step flop ratio
(scale_cpu_sample), the raw rate stays in cpu_baseline.sample_value.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (generator kwargs for the GPU arm, kwargs for the bounded CPU sample, sample note)
    "C1": ("c1_random_qp", {}, {}, "full C1 instance"),
    "C2": ("c2_portfolio", {}, {"n": 20000}, "same generator at n=2e4 (1/5 of the assets)"),
    "C3": ("c3_socp", {}, {"n": 100000, "ncones": 2000}, "same generator at n=1e5, 2000 cones (1/5)"),
    "C4": ("c4_sdp", {}, {"ncones": 20, "side": 30, "n": 3000, "vars_per_cone": 200},
           "same generator, 20 cones of side 30"),
    "C4r": ("c4_sdp", {"ncones": 40, "side": 40, "n": 8000, "vars_per_cone": 300},
            {"ncones": 20, "side": 30, "n": 3000, "vars_per_cone": 200},
            "reduced C4 (40 PSD cones of side 40); CPU sample: 20 cones of side 30"),
    "C5": ("c5_block_angular", {}, {"nblocks": 2, "nlink": 64},
           "2 of the 64 diagonal blocks with proportionally fewer linking rows (same generator, "
           "nblocks=2, nlink=64): 1/32 of the full instance; CPU time per iteration grows at least "
           "linearly in the number of blocks"),
}


def describe(name, P, A, ks):
    info = ks.ldl.info() if hasattr(ks, "ldl") and hasattr(ks.ldl, "info") else None
    return dict(workload=name, n=int(A.shape[1]), m=int(A.shape[0]), N=int(ks.KKT.shape[0]),
                nnzK=int(ks.KKT.nnz), nnzL=int(info.nnzL) if info else int(ks.ldl.nnzL))


REPLAY_ITERS = 2      # IP iterations run to record realistic cone states / right-hand sides
SOLVES_PER_STEP = 6.0  # 3 right-hand sides, each with one refinement round (observed on C1..C5)


def kkt_step_work(factor_flops, nnzL, N):
    """Algorithmic flops of one step: numeric LDL' = sum_j l_j^2, plus
    SOLVES_PER_STEP triangular solves of 4 nnzL + N each."""
    return float(factor_flops) + SOLVES_PER_STEP * (4.0 * float(nnzL) + float(N))


def full_size_work(KKT):
    """Work of one step on the full workload under the CPU path's AMD-class ordering (host-only
    symbolic analysis of this repo's library; no numerics)."""
    from clarabel_jl_b200 import lib as cblib
    st = cblib.Symbolic(KKT, ordering=0).stats
    return kkt_step_work(st["flops"], st["nnzL"], st["N"]), dict(N=int(st["N"]), nnzL=int(st["nnzL"]), factor_flops=float(st["flops"]))


def scale_cpu_sample(name, rate, ms, sdesc, note, full_KKT_fn):
    """The CPU arm runs a bounded sample; its rate is converted to the metric's unit (iterations
    of the FULL workload per second) by the ratio of algorithmic work per step.  This assumes the
    scalar CPU code keeps its sample flop rate on the 30-100x larger factor (optimistic for the CPU:
    the sample's L fits in cache, the full one does not)."""
    gen, kw, skw, _ = WORKLOADS[name]
    w_s = kkt_step_work(sdesc["factor_flops"], sdesc["nnzL"], sdesc["N"])
    if kw == skw:
        return rate, dict(value=rate, unit="it/s", cores=1, kind="port", sample=note, sample_config=sdesc,
                          sample_value=rate, sample_ms_per_step=ms, work_ratio=1.0)
    w_f, fdesc = full_size_work(full_KKT_fn())
    ratio = w_s / w_f
    cb_ = dict(value=rate * ratio, unit="it/s", cores=1, kind="port",
               sample=note + "; value = measured sample rate x (flops of one sample step / flops of one "
               "full-size step), flops = sum_j l_j^2 + 6 (4 nnzL + N) under the AMD-class ordering",
               sample_config=sdesc, sample_value=rate, sample_ms_per_step=ms, work_ratio=ratio,
               full_size=fdesc)
    return rate * ratio, cb_


class Recorder:
    """Wraps a KKT solver and records the inputs of every boundary call of an IP run."""
    def __init__(self, ks, cones):
        self.ks, self.cones = ks, cones
        self.steps = []          # list of dict(state=..., rhs=[(rx, rz), ...])
        self._u, self._s, self._r = ks.update, ks.solve, ks.setrhs
        ks.update, ks.solve, ks.setrhs = self.update, self.solve, self.setrhs
        self._rhs = None

    def update(self, cones):
        st = {k: np.array(v, copy=True) for k, v in cones.export_state().items()}
        self.steps.append(dict(state=st, rhs=[]))
        return self._u(cones)

    def setrhs(self, rx, rz):
        self._rhs = (np.array(rx, copy=True), np.array(rz, copy=True))
        return self._r(rx, rz)

    def solve(self, lx, lz):
        if self.steps:
            self.steps[-1]["rhs"].append(self._rhs)
        return self._s(lx, lz)

    def detach(self):
        self.ks.update, self.ks.solve, self.ks.setrhs = self._u, self._s, self._r


class FakeCones:
    def __init__(self, st):
        self.st = st

    def export_state(self):
        return self.st


def sample_clocks(stop, out):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                os.environ.get("LOCAL_RANK", "0")], capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split("\n")[0].split(",")]
            out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    if not samples:
        return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
    sm = [float(s[1]) for s in samples if len(s) > 2]
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        for k, nm in enumerate(names):
            if len(s) > 5 + k and s[5 + k].lower().startswith("active"):
                reasons.add(nm)
    return dict(sm_mhz=float(np.median(sm)) if sm else None,
                sm_max_mhz=float(samples[0][2]) if len(samples[0]) > 2 else None,
                reasons=sorted(reasons))


def cpu_sample(name, steps, warmup):
    """Times the CPU oracle (QDLDL-algorithm restatement, 1 thread) on the bounded sample."""
    import clarabel_jl_b200 as cb
    from clarabel_jl_b200 import problems
    from oracle.kktsolver_oracle import OracleDirectLDLKKTSolver
    cb.register_kktsolver("qdldl", OracleDirectLDLKKTSolver)
    gen, _, skw, note = WORKLOADS[name]
    P, q, A, b, K = getattr(problems, gen)(**skw)
    s = cb.Solver(P, q, A, b, K, cb.Settings(direct_solve_method="qdldl"))
    ks = s.kktsystem.kktsolver
    rec = Recorder(ks, s.cones)
    s.solve(max_iter=REPLAY_ITERS)
    rec.detach()
    replay = [st for st in rec.steps[1:] if len(st["rhs"]) == 3] or rec.steps[-1:]
    lx, lz = np.zeros(s.data.n), np.zeros(s.data.m)

    def step(i):
        st = replay[i % len(replay)]
        ok = ks.update(FakeOracleCones(s.cones, st["state"]))
        for rx, rz in st["rhs"]:
            ks.setrhs(rx, rz); ok &= ks.solve(lx, lz)
        return ok
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = time.perf_counter() - t0
    desc = describe(name + "-sample", P, A, ks)
    desc["factor_flops"] = float(ks.ldl.sum_lnz_sq)
    return steps / dt, dt / steps * 1e3, desc, note


class FakeOracleCones:
    """A CompositeCone whose scaling state is overwritten by a recorded one (CPU oracle replay)."""
    def __new__(cls, cones, st):
        cones.w[:] = st["w"]; cones.soc_eta[:] = st["soc_eta"]; cones.soc_d[:] = st["soc_d"]
        cones.soc_u[:] = st["soc_u"]; cones.soc_v[:] = st["soc_v"]
        off = 0
        psd = [i for i, t in enumerate(cones.types) if t == 3]
        for g in cones.psd_groups:
            pass
        if len(st["psd_R"]):
            pos = {}
            for g in cones.psd_groups:
                for j, ci in enumerate(g["cones"]):
                    pos[ci] = (g, j)
            for ci in psd:
                g, j = pos[ci]; n = g["n"]
                g["R"][j] = st["psd_R"][off:off + n * n].reshape(n, n, order="F"); off += n * n
        return cones


def _watchdog(seconds):
    """A hung collective must not hang the caller: dump stacks and exit non-zero."""
    import faulthandler
    faulthandler.dump_traceback_later(seconds, exit=True)


def main():
    _watchdog(int(os.environ.get("CB200_BENCH_TIMEOUT", "840")))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CB200_WORKLOAD", "C5"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    metric = "IP-iterations/sec (KKT assemble+factor+solve)"

    if args.impl == "reference":
        if rank != 0:
            return
        rate, ms, desc, note = cpu_sample(name, max(1, args.steps), max(0, min(args.warmup, 1)))

        def full_kkt():
            import clarabel_jl_b200 as cb
            from clarabel_jl_b200 import problems, kkt_assembly as ka
            gen, kw, _, _ = WORKLOADS[name]
            P, q, A, b, K = getattr(problems, gen)(**kw)
            data = cb.problemdata.ProblemData(P, q, A, b, K, cb.Settings())
            return ka.assemble_kkt_matrix(data.P, data.A, cb.CompositeCone(data.cones))[0]
        v, cbl = scale_cpu_sample(name, rate, ms, desc, note, full_kkt)
        line = dict(metric=metric, value=v, unit="it/s", n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=1e3 / v, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                    config=dict(workload=name, sample=cbl["sample"], **{k: v_ for k, v_ in cbl.get("full_size", desc).items()
                                                                        if k in ("N", "nnzL")}),
                    cpu_baseline=cbl,
                    e2e=dict(value=v, unit="it/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    import clarabel_jl_b200 as cb
    from clarabel_jl_b200 import problems, lib as cblib
    gen, kw, _, _ = WORKLOADS[name]
    P, q, A, b, K = getattr(problems, gen)(**kw)
    st = cb.Settings(direct_solve_method="b200")
    t0 = time.perf_counter()
    solver = cb.Solver(P, q, A, b, K, st)
    t_setup = time.perf_counter() - t0
    ks = solver.kktsystem.kktsolver
    cblib.make_settings  # noqa
    rec = Recorder(ks, solver.cones)
    solver.solve(max_iter=REPLAY_ITERS)
    rec.detach()
    replay = [s_ for s_ in rec.steps[1:] if len(s_["rhs"]) == 3] or rec.steps[-1:]
    n, m = solver.data.n, solver.data.m

    def pinned(a):
        """copy of `a` in page-locked host memory (the e2e copies are then true DMA transfers)"""
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy(); v[...] = a
        pinned.keep.append(t)
        return v
    pinned.keep = []
    for s_ in replay:
        s_["state"] = {k: pinned(v) for k, v in s_["state"].items()}
        s_["rhs"] = [(pinned(rx), pinned(rz)) for rx, rz in s_["rhs"]]
    lx, lz = pinned(np.zeros(n)), pinned(np.zeros(m))
    stream = torch.cuda.ExternalStream(ks.ldl.stream_ptr(), device=torch.device("cuda", local))

    def step(i):
        s_ = replay[i % len(replay)]
        ok = ks.update(FakeCones(s_["state"]))
        for rx, rz in s_["rhs"]:
            ks.setrhs(rx, rz); ok &= ks.solve(lx, lz)
        return ok

    def timed(nsteps, offset):
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(nsteps):
            step(offset + i)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    state_bytes = sum(v.nbytes for v in replay[0]["state"].values())
    h2d = state_bytes + 3 * (n + m) * 8
    d2h = 3 * (n + m) * 8
    # ---- e2e (host buffers through the C-ABI)
    for i in range(args.warmup):
        step(i)
    stop = threading.Event(); samples = []
    th = threading.Thread(target=sample_clocks, args=(stop, samples), daemon=True)
    th.start()
    ks.ldl.reset_timers()
    ms_e2e = timed(args.steps, args.warmup)
    # ---- device-resident
    ks.ldl.set_resident(True)
    for i in range(args.warmup):
        step(i)
    ks.ldl.reset_timers()
    ms_res = timed(args.steps, args.warmup)
    tm = ks.ldl.timers()
    stop.set(); th.join(timeout=2)
    ks.ldl.set_resident(False)
    # ---- per-kernel-class timing of the factorisation (events around each launch group; graph
    # replay is off for these extra steps, they are not part of the timed regions above).  In
    # multi-GPU mode EVERY rank must execute them (they contain collectives).
    ks.ldl.set_detail(True); ks.ldl.set_resident(True)
    step(0)
    ks.ldl.reset_timers()
    for i in range(2):
        step(i)
    td = ks.ldl.timers()
    ks.ldl.set_detail(False); ks.ldl.set_resident(False)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    desc = describe(name, P, A, ks)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
    nfac, nsol = max(1, tm["nfactor"]), max(1, tm["nsolve"])
    N, nnzL, nnzK = desc["N"], desc["nnzL"], desc["nnzK"]
    stats = ks.ldl.stats()
    # ---- FP64 compute calibration on this box (no FP64 figure in MEASURED_PEAKS.json): cuBLAS DGEMM
    fp64_peak = None
    try:
        a_ = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        b_ = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        for _ in range(2):
            torch.matmul(a_, b_)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a_, b_); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fp64_peak = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
        del a_, b_
    except Exception:
        pass
    nfd = max(1, td["nfactor"])
    b_solve = 16.0 * nnzL + 48.0 * N
    b_spmv = 8.0 * nnzK + 24.0 * N
    t_solve = tm["solve_ms"] / nsol * 1e-3
    t_fac = tm["factor_ms"] / nfac * 1e-3
    t_spmv = tm["spmv_ms"] / nsol * 1e-3
    t_schur = td["schur_ms"] / nfd * 1e-3
    kernels = dict(
        triangular_solve_sweeps=dict(bound="hbm", ms_per_call=t_solve * 1e3, calls_per_step=nsol / args.steps,
                                     achieved=b_solve / t_solve / 1e9, unit="GB/s", peak=hbm_peak,
                                     algorithmic_bytes=b_solve),
        spmv_residual=dict(bound="hbm", ms_per_call=t_spmv * 1e3, calls_per_step=nsol / args.steps,
                           achieved=b_spmv / max(t_spmv, 1e-12) / 1e9, unit="GB/s", peak=hbm_peak,
                           algorithmic_bytes=b_spmv),
        factor_total=dict(bound="fp64", ms_per_call=t_fac * 1e3, calls_per_step=nfac / args.steps,
                          achieved=stats["flops"] / t_fac / 1e12, unit="TFLOP/s", peak=fp64_peak,
                          algorithmic_flops=stats["flops"]),
        k_schur_large=dict(bound="fp64", ms_per_call=t_schur * 1e3, calls_per_step=1.0,
                           achieved=(stats["schur_flops"] / t_schur / 1e12) if t_schur > 0 else None,
                           unit="TFLOP/s", peak=fp64_peak, algorithmic_flops=stats["schur_flops"]),
        factor_pivot_blocks=dict(ms_per_call=td["panel_ms"] / nfd), factor_small_fronts=dict(ms_per_call=td["small_ms"] / nfd),
        factor_assembly=dict(ms_per_call=td["asm_ms"] / nfd))
    for k in kernels.values():
        if k.get("peak") and k.get("achieved") is not None:
            k["frac"] = k["achieved"] / k["peak"]
    # dominant kernel of the step: the largest (time per call x calls per step) among the measured ones
    cand = {"triangular_solve_sweeps": t_solve * nsol / args.steps, "k_schur_large": t_schur,
            "spmv_residual": t_spmv * nsol / args.steps}
    dom = max(cand, key=cand.get)
    kd = kernels[dom]
    roofline = dict(bound=kd["bound"], kernel=dom, achieved=kd["achieved"], peak=kd["peak"], unit=kd["unit"],
                    frac=kd.get("frac"), traffic=None,
                    peak_source=(peak_src if kd["bound"] == "hbm" else
                                 "on-box cuBLAS DGEMM 4096^3 (float64) - MEASURED_PEAKS.json has no FP64 figure"),
                    note="bound 'fp64' = FP64 FMA pipe (tcgen05 has no f64 kind; see DESIGN.md section 4)",
                    kernels=kernels)
    line = dict(metric=metric, value=args.steps / (ms_res * 1e-3), unit="it/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms_res / args.steps,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", impl="b200",
                config=dict(desc, l2="inputs larger than L2 (no flush)", setup_s=t_setup,
                            replayed_ip_iterations=len(replay), ordering="nd+amd"),
                e2e=dict(value=args.steps / (ms_e2e * 1e-3), unit="it/s", ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
                gpu_launches=int(tm["nlaunch"]), clocks=clocks_summary(samples), roofline=roofline)
    if not args.no_cpu_baseline and world == 1:      # the CPU arm is timed at N = 1 only
        try:
            rate, ms, sdesc, note = cpu_sample(name, 1, 1)
            line["cpu_baseline"] = scale_cpu_sample(name, rate, ms, sdesc, note, lambda: ks.KKT)[1]
        except Exception as e:                      # the GPU measurement above must still be reported
            line["cpu_baseline"] = dict(value=None, unit="it/s", cores=1, kind="port",
                                        sample=f"CPU leg failed: {type(e).__name__}: {e}")
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
