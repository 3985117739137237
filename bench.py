#!/usr/bin/env python
"""bench.py -- DoF-steps/s of the wave-equation time-stepping hot path on B200, with the SpMV
roofline and the CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one pass of the reference's time loop body (src/WaveNewmark.cpp:424-440): assemble_rhs,
the Dirichlet values, the Jacobi-PCG solve, the Newmark update and the norms.  Default workload at
N=1: BASELINE.json configs[1] (standing-mode-wsol.json, Newmark beta=1/4 gamma=1/2, Nel=1024, R=1).
At N>1 the default is the weak-scaled version of the same workload (Nel = 1024 x 1024*N on
[0,1]x[0,N], one strip of 1024 quad rows per GPU).  Other workloads: see WORKLOADS.

Output: one JSON line (rank 0).  Timing: CUDA events on the context's stream around every step,
L2 flushed between steps, max over ranks.  The oracle (oracle/) is used only for the cpu_baseline
leg and for --impl reference; nothing here reads /root/reference."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent

WORKLOADS = {
    # name: (problem, scheme, overrides, scaling, weak-scale-in-y)
    "c2-standing-newmark-1024-p1": ("standing-mode-wsol", "newmark",
                                    dict(Nel="1024", R="1", Dt="0.01", Beta="0.25", Gamma="0.5"), "weak", True),
    "c3-gaussian-explicit-4096-p2": ("gaussian-pulse", "newmark",
                                     dict(Nel="4096", R="2", Dt="3.8e-5", Beta="0.0", Gamma="0.5"), "strong", False),
    "c4-ricker-be-4096-p2": ("ricker-wavelet", "theta", dict(Nel="4096", R="2", Theta="1.0"), "strong", False),
    "c5-traveling-newmark-2896-p2": ("traveling-square-bump", "newmark",
                                     dict(Nel="2896", R="2", Geometry="[0.0, 3.0] x [0.0, 3.0]",
                                          C={"Function constants": "", "Variable names": "x, y, t",
                                             "Function expression": "1.0 + 0.25*sin(2*pi*x/3)*sin(2*pi*y/3)"}),
                                     "weak", True),
    # the reference's own published strong-scaling configuration (BASELINE.md section 1:
    # report/sections/8_Scalability.tex:9-18): Nel=640, R=1, Dt=8e-5, Newmark 1/4, 1/2
    "published-newmark-640-p1": ("standing-mode-wsol", "newmark", dict(Nel="640", R="1", Dt="8e-5"), "strong", False),
    # BASELINE configs[4] at its full size (268 M DoFs, 3.09 G nnz): needs >= 2 GPUs
    "c5-traveling-newmark-8192-p2": ("traveling-square-bump", "newmark",
                                     dict(Nel="8192", R="2", Geometry="[0.0, 3.0] x [0.0, 3.0]",
                                          C={"Function constants": "", "Variable names": "x, y, t",
                                             "Function expression": "1.0 + 0.25*sin(2*pi*x/3)*sin(2*pi*y/3)"}),
                                     "strong", False),
    "newmark-4096-p2": ("standing-mode-wsol", "newmark", dict(Nel="4096", R="2", Dt="0.001"), "strong", False),
    "newmark-2048-p2": ("standing-mode-wsol", "newmark", dict(Nel="2048", R="2", Dt="0.002"), "strong", False),
}
DEFAULT = "c2-standing-newmark-1024-p1"
# BASELINE.md publishes a number only for this workload: 410 881 DoFs x 625 steps / 296.3 s whole-process
# wall time on one Xeon Gold 6238R core (AMG-CG); 16 ranks: 9.31 M, 32 ranks: 12.8 M DoF-steps/s
PUBLISHED = {"published-newmark-640-p1": 0.867e6}


def make_params(workload, n_gpus):
    from wavegpu.problems import problem

    name, scheme, over, scaling, weak_y = WORKLOADS[workload]
    over = dict(over)
    if os.environ.get("WAVE_BENCH_NEL"):  # debugging aid: same problem on another mesh
        over["Nel"] = os.environ["WAVE_BENCH_NEL"]
    p = problem(name, **over)
    if scaling == "weak" and n_gpus > 1 and weak_y:
        from wavegpu.api import parse_geometry, parse_nel

        nx, ny = parse_nel(p["Nel"])
        x0, x1, y0, y1 = parse_geometry(p["Geometry"])
        p["Nel"] = f"{nx}, {ny * n_gpus}"
        p["Geometry"] = f"[{x0}, {x1}] x [{y0}, {y0 + (y1 - y0) * n_gpus}]"
    return p, scheme, scaling


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def hbm_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def time_oracle(params, scheme, budget_s, max_steps, warmup=1, precond=0):
    """The CPU path: oracle/wave_oracle.c (C restatement of the reference, OpenMP over all host
    threads) stepping the same workload for a bounded number of steps."""
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU path is meant to use every host core
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    from oracle import oracle as O

    O.lib().oracle_set_num_threads(ncores)  # in case the OpenMP runtime was initialised before
    t0 = time.time()
    o = O.Oracle.from_params(params)
    o.set_cg(precond=precond)
    dt = float(params["Dt"])
    if scheme == "newmark":
        o.newmark_init(dt, float(params["Beta"]), float(params["Gamma"]))
        step = o.newmark_step
    else:
        o.theta_init(dt, float(params["Theta"]))
        step = o.theta_step
    setup_s = time.time() - t0
    for _ in range(warmup):
        step()
    times, its = [], 0
    t_begin = time.time()
    while len(times) < max_steps and (time.time() - t_begin) < budget_s:
        t1 = time.time()
        step()
        times.append(time.time() - t1)
        its += sum(o.iterations())
    total = sum(times)
    return {"n": o.n, "steps": len(times), "seconds": total, "setup_s": setup_s,
            "cg_its_per_step": its / max(len(times), 1), "threads": O.lib().oracle_num_threads()}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The deal.II/Trilinos
    binaries cannot be built in this image, so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload or DEFAULT
    params, scheme, scaling = make_params(workload, 1)
    r = time_oracle(params, scheme, budget_s=150.0, max_steps=args.steps, warmup=min(args.warmup, 2))
    value = r["n"] * r["steps"] / r["seconds"]
    line = {
        "impl": "reference", "metric": "dof_steps_per_sec", "value": value, "unit": "DoF-steps/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 2),
        "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "n_dofs": r["n"], "cg_its_per_step": r["cg_its_per_step"],
                   "preconditioner": "jacobi"},
        "cpu_baseline": {"value": value, "unit": "DoF-steps/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{r['steps']} time steps of the same workload (oracle/wave_oracle.c, OpenMP)"},
        "e2e": {"value": value, "unit": "DoF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precond", default="jacobi", choices=["jacobi", "mg"],
                    help="CG preconditioner: jacobi (north-star default) or the multigrid V-cycle (single GPU)")
    ap.add_argument("--cg", default="three-kernel", choices=["three-kernel", "fused"],
                    help="CG iteration: the measured three-kernel path (default) or the experimental cooperative "
                         "kernel K6f (csrc/cg_fused.cu, WAVE_CG_FUSED=1; falls back when the problem does not fit)")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-scale-probe", action="store_true",
                    help="skip the SpMV / CG-iteration roofline probe on matrices larger than L2")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.cg == "fused":
        os.environ["WAVE_CG_FUSED"] = "1"  # read by wave_setup of every context created below

    import numpy as np
    import torch

    from wavegpu import WaveSolver, api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(api.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    n_gpus = world
    workload = args.workload or DEFAULT
    params, scheme, scaling = make_params(workload, n_gpus)
    # a real (non-default) stream: the library runs on the stream it is given, so the events below are
    # recorded on the very stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    W = max(args.warmup, 3)
    K = args.steps

    t_setup0 = time.time()
    cg_opts = dict(precond=2) if args.precond == "mg" else None
    g = WaveSolver(params, scheme, rank=rank, nranks=world, nccl_id=nccl_id, device=local_rank,
                   stream=stream.cuda_stream, cg=cg_opts)
    g.init()
    setup_s = time.time() - t_setup0
    cg_fused = g.cg_fused_active()
    n = g.n
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up through the timed and the bracketed pass
    # (the timed region alone lasts only tens of milliseconds at 100 ms sampling)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        g.step()
    # ---- timed region: K steps, device timed on the context stream -------------------------------
    g.cg_stats(reset=True)
    launches0 = g.launch_count()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    its_total = 0
    wall0 = time.time()
    for k in range(K):
        if flush is not None:
            flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        its, _ = g.step()
        ev[k][1].record(stream)
        its_total += its[0] + its[1]
    barrier()
    wall_s = time.time() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    launches = g.launch_count() - launches0
    cgs = g.cg_stats()
    # second pass over further steps of the same run with every CG SpMV launch bracketed by events on
    # the context stream (kept out of the pass above: the brackets cost ~1 us per launch)
    g.spmv_timing(True)
    Kr = max(3, min(K, 10))
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kr)]
    for k in range(Kr):
        if flush is not None:
            flush.fill_(k & 0xFF)
        ev2[k][0].record(stream)
        g.step()
        ev2[k][1].record(stream)
    barrier()
    spmv_launches, spmv_ms = g.spmv_timing(False)
    bracketed_ms = float(sum(a.elapsed_time(b) for a, b in ev2))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = n * K / (total_ms * 1e-3)

    # ---- e2e: the stateless host-buffer entry point, H2D + D2H of the state every step -------------
    e2e = None
    if K > 0:
        def pinned(vec):
            t = torch.empty(n, dtype=torch.float64).pin_memory()
            t.numpy()[g.row0:g.row0 + g.nown] = vec
            return t.numpy()

        # every rank keeps (global-length) host arrays and moves its own rows each step
        lo, hi = g.row0, g.row0 + g.nown
        if world == 1:
            u0_, v0_, a0_ = g.vector(api.VEC_U), g.vector(api.VEC_V), g.vector(api.VEC_A) if scheme == "newmark" else None
        else:
            u0_, v0_ = g.vector_owned(api.VEC_U), g.vector_owned(api.VEC_V)
            a0_ = g.vector_owned(api.VEC_A) if scheme == "newmark" else None
        u, v = pinned(u0_), pinned(v0_)
        a = pinned(a0_) if scheme == "newmark" else None
        nvec = 3 if scheme == "newmark" else 2
        for _ in range(2):
            g.step_host(u, v, a)
        barrier()
        Ke = max(3, min(K, 10))
        t0 = time.time()
        for _ in range(Ke):
            g.step_host(u, v, a)
        barrier()
        e2e_s = time.time() - t0
        if world > 1:
            tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        e2e = {"value": n * Ke / e2e_s, "unit": "DoF-steps/s", "h2d_bytes_per_step": nvec * 8 * n,
               "d2h_bytes_per_step": nvec * 8 * n + 16 * world, "steps": Ke,
               "api": "wave_step_host (pinned host u, v, a in; u, v, a, norms out; every rank moves its own rows)"}

    # ---- roofline of the dominant kernel: the CG SpMV, timed live inside the steps above ------------
    peak, peak_src = hbm_peak()
    nnz = g.nnz_local
    alg_bytes = 12.0 * nnz + 20.0 * g.nown
    avg_ms = spmv_ms / spmv_launches if spmv_launches else float("nan")
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if spmv_launches else float("nan")
    ms_fl, _ = g.bench_spmv(api.MAT_SYS1, reps=20, flush_l2=True)
    ms_hot, _ = g.bench_spmv(api.MAT_SYS1, reps=50, flush_l2=False)
    traffic = None
    tf = ROOT / "profiles" / "spmv_traffic.json"
    if tf.exists():  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.loads(tf.read_text()).get(workload, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_spmv<1,false> (CG A*d with fused d.Ad)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms_in_step": avg_ms, "launches_timed": spmv_launches,
                "share_of_step_time": spmv_ms / bracketed_ms if bracketed_ms else None,
                "timed_over": f"{Kr} further steps of the same run, every CG SpMV launch bracketed by CUDA events",
                "l2_flushed_single_launch": {"ms": ms_fl, "GB/s": alg_bytes / (ms_fl * 1e-3) / 1e9,
                                             "frac": alg_bytes / (ms_fl * 1e-3) / 1e9 / peak},
                "back_to_back": {"ms": ms_hot, "GB/s": alg_bytes / (ms_hot * 1e-3) / 1e9},
                "note": "working set vs 126 MB L2: matrix %.0f MB + vectors; inside a CG solve the matrix is "
                        "re-read every iteration, so small workloads run partly from L2" % (12.0 * nnz / 1e6)}
    if cg_fused:
        # K6f: no per-SpMV launches to bracket; the figure is the whole solve (start residual, every
        # iteration, the host's read-back) over its iterations, against 12 nnz + 26 n bytes per iteration
        fb = 12.0 * nnz + 26.0 * g.nown
        ms_it = cgs["ms_total"] / cgs["iterations"] if cgs["iterations"] else float("nan")
        roofline.update({"kernel": "k_cg_fused (one cooperative kernel per solve: SpMV from staged d, updates on chip)",
                         "algorithmic_bytes_per_launch": fb, "avg_launch_ms_in_step": ms_it,
                         "achieved": fb / (ms_it * 1e-3) / 1e9, "frac": fb / (ms_it * 1e-3) / 1e9 / peak,
                         "launches_timed": cgs["iterations"],
                         "share_of_step_time": cgs["ms_total"] / total_ms if total_ms else None,
                         "timed_over": "all CG solves of the timed steps: device ms per solve / iterations",
                         "traffic": None})
    if rank != 0:
        g.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the same kernels on matrices far larger than the 126 MB L2 (the bench workload's 88 MB matrix
    # makes a 25 us kernel: launch ramp and tail weigh on its fraction) -------------------------------
    at_scale = None
    if world == 1 and not args.no_scale_probe:
        from wavegpu.problems import problem as _problem

        at_scale = []
        for nel, r in (("4096", 1), ("2048", 2)):
            gs = WaveSolver(_problem("standing-mode-wsol", Nel=nel, R=r, Dt="0.002"), "newmark",
                            stream=stream.cuda_stream)
            gs.init()
            ms_s, by_s = gs.bench_spmv(api.MAT_SYS1, reps=10, flush_l2=True)
            ms_c, by_c = gs.bench_cg_iter(api.MAT_SYS1, reps=2)
            at_scale.append({"Nel": nel, "R": r, "n_dofs": gs.n, "nnz": gs.nnz_local,
                             "spmv": {"ms": ms_s, "GB/s": by_s / ms_s / 1e6, "frac": by_s / ms_s / 1e6 / peak,
                                      "algorithmic_bytes": by_s, "l2": "flushed before every launch"},
                             "cg_iteration": {"ms": ms_c, "GB/s": by_c / ms_c / 1e6, "frac": by_c / ms_c / 1e6 / peak,
                                              "algorithmic_bytes": by_c}})
            gs.close()
    # ---- the same workload with the multigrid V-cycle preconditioner (single GPU; the default `value`
    # keeps the north star's Jacobi so that every N runs the same algorithm) ---------------------------
    multigrid = None
    if world == 1 and args.precond == "jacobi" and not args.no_scale_probe:
        gm = WaveSolver(params, scheme, stream=stream.cuda_stream, cg=dict(precond=2))
        gm.init()
        for _ in range(W):
            gm.step()
        torch.cuda.synchronize()
        evm = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        its_m = 0
        for k in range(K):
            if flush is not None:
                flush.fill_(k & 0xFF)
            evm[k][0].record(stream)
            it_m, _ = gm.step()
            evm[k][1].record(stream)
            its_m += it_m[0] + it_m[1]
        torch.cuda.synchronize()
        ms_m = float(sum(a.elapsed_time(b) for a, b in evm))
        multigrid = {"value": n * K / (ms_m * 1e-3), "unit": "DoF-steps/s", "ms_per_step": ms_m / K,
                     "cg_its_per_step": its_m / K,
                     "preconditioner": "geometric multigrid V(2,2), damped Jacobi smoothing (WAVE_PRECOND_MG)"}
        gm.close()
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_oracle(params, scheme, budget_s=20.0, max_steps=10, precond=2 if args.precond == "mg" else 0)
        cpu = {"value": r["n"] * r["steps"] / r["seconds"], "unit": "DoF-steps/s", "cores": r["threads"],
               "kind": "port", "cg_its_per_step": r["cg_its_per_step"],
               "sample": f"{r['steps']} time steps of the same workload on the host "
                         f"(oracle/wave_oracle.c, OpenMP, {r['threads']} threads)"}

    line = {
        "metric": "dof_steps_per_sec", "value": value, "unit": "DoF-steps/s", "n_gpus": n_gpus, "steps": K,
        "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": (value / PUBLISHED[workload]) if workload in PUBLISHED else None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "problem": WORKLOADS[workload][0], "scheme": scheme,
                   "Nel": params["Nel"], "R": params["R"], "Dt": params["Dt"], "n_dofs": n,
                   "nnz_per_gpu": nnz, "cg_its_per_step": its_total / K, "preconditioner": args.precond,
                   "cg_path": "fused (K6f)" if cg_fused else "three-kernel",
                   "cg_stop": "ReductionControl(10000, 1e-12, 1e-6)", "parallelism": f"strips{n_gpus}",
                   "l2": "flushed between steps (256 MiB write)" if flush is not None else "not flushed",
                   "setup_s": setup_s, "wall_s_timed_region": wall_s},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        "roofline_at_scale": at_scale, "multigrid": multigrid, "cpu_baseline": cpu,
        "cg": {"solves": cgs["solves"], "iterations": cgs["iterations"], "ms_total": cgs["ms_total"],
               "ms_per_iteration": cgs["ms_total"] / max(cgs["iterations"], 1)},
        "step_ms": {"min": min(step_ms), "max": max(step_ms)},
    }
    print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
