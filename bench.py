#!/usr/bin/env python
"""bench.py -- DoF-steps/s of the wave-equation time-stepping hot path on B200, with the SpMV
roofline and the CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one pass of the reference's time loop body (src/WaveNewmark.cpp:424-440): assemble_rhs,
the Dirichlet values, the Jacobi-PCG solve, the Newmark update and the norms.  Default workload at every
N: the north star's 64 M-DoF P2 Newmark run (standing-mode-wsol.json, Newmark beta=1/4 gamma=1/2,
Nel=4096, R=2, 67 125 249 DoFs) -- strong scaling, the same global problem on 1, 2, 4, 8 GPUs.  At N=1 the
line also carries BASELINE.json's other single-GPU configs as `extras` (c2, c3, c4, and the default
workload with the multigrid preconditioner).  Other workloads: see WORKLOADS.

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
DEFAULT = "newmark-4096-p2"  # the north star's 64 M-DoF P2 Newmark run; strong scaling at every N
EXTRAS = ["c2-standing-newmark-1024-p1", "c3-gaussian-explicit-4096-p2", "c4-ricker-be-4096-p2"]
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


def cpu_sample_params(params):
    """Bounded sample of a workload for the CPU arm: the same mesh widths (dx, dy), time step and scheme on
    a strip of fewer quad rows, about 4 M DoFs (the full 67 M-DoF workloads need ~140 s of set-up and ~20 s
    per step on the host).  The functions are squeezed into the strip (y -> y0 + S (y - y0), S = ny / ny_s) so
    that initial and boundary data stay compatible and the CG solves see the same kind of right-hand sides
    (iterations per step are reported next to the figure).  Small workloads are taken whole."""
    import copy
    import re

    from wavegpu.api import parse_geometry, parse_nel

    nx, ny = parse_nel(params["Nel"])
    r = int(params["R"])
    per_row = r * (r * nx + 1)  # DoFs added per quad row
    ny_s = max(8, int(round(4.2e6 / per_row)))
    if ny_s >= ny:
        return dict(params), "the whole workload"
    x0, x1, y0, y1 = parse_geometry(params["Geometry"])
    p = copy.deepcopy(params)
    p["Nel"] = f"{nx}, {ny_s}"
    p["Geometry"] = f"[{x0}, {x1}] x [{y0}, {y0 + (y1 - y0) * ny_s / ny!r}]"
    scale = ny / ny_s
    for blk in ("C", "F", "U0", "V0", "G", "DGDT", "Solution"):
        if blk in p and isinstance(p[blk], dict) and p[blk].get("Function expression"):
            p[blk]["Function expression"] = re.sub(r"\by\b", f"({y0!r} + {scale!r}*(y - {y0!r}))",
                                                   p[blk]["Function expression"])
    return p, (f"strip of {ny_s} of the workload's {ny} quad rows (same dx, dy, Dt; Nel = {nx} x {ny_s}; "
               f"functions squeezed in y by {scale:g})")


def time_oracle(params, scheme, budget_s, max_steps, warmup=1, precond=0):
    """The CPU path: oracle/wave_oracle.c (C restatement of the reference, OpenMP over all host
    threads) stepping the sample for a bounded number of steps."""
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


def workload_config(workload, params, scheme, n_dofs, precond="jacobi"):
    """The part of `config` that names the workload: identical for the GPU arm and the reference arm."""
    return {"workload": workload, "problem": WORKLOADS[workload][0], "scheme": scheme, "Nel": params["Nel"],
            "R": str(params["R"]), "Dt": params["Dt"], "n_dofs": int(n_dofs), "preconditioner": precond,
            "cg_stop": "ReductionControl(10000, 1e-12, 1e-6)",
            "timing": "GPU arm: CUDA events around every step on the context's stream, L2 flushed between steps "
                      "(256 MiB write; the working set is far larger than L2 anyway); reference arm: host wall "
                      "clock around every step"}


def full_n_dofs(params):
    from wavegpu.api import parse_nel

    nx, ny = parse_nel(params["Nel"])
    r = int(params["R"])
    return (r * nx + 1) * (r * ny + 1)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The deal.II/Trilinos
    binaries cannot be built in this image, so this is the oracle port, on all
    host threads, on a bounded sample of the GPU arm's workload (same config, metric and unit)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload or DEFAULT
    params, scheme, scaling = make_params(workload, args.gpus)
    sample, what = cpu_sample_params(params)
    W = max(args.warmup, 3)
    r = time_oracle(sample, scheme, budget_s=150.0, max_steps=args.steps, warmup=W)
    value = r["n"] * r["steps"] / r["seconds"]
    sample_txt = (f"{what}: {r['n']} DoFs, {r['steps']} time steps after {W} warm-up steps "
                  f"(oracle/wave_oracle.c, OpenMP, {r['threads']} threads); DoF-steps/s is a rate, "
                  "so the sample's figure stands for the workload")
    line = {
        "impl": "reference", "metric": "dof_steps_per_sec", "value": value, "unit": "DoF-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": W,
        "ms_per_step": 1e3 * r["seconds"] / r["steps"] * (full_n_dofs(params) / r["n"]),
        "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(workload, params, scheme, full_n_dofs(params)),
        "run": {"steps_timed": r["steps"], "sample_n_dofs": r["n"], "sample_ms_per_step": 1e3 * r["seconds"] / r["steps"],
                "cg_its_per_step": r["cg_its_per_step"], "setup_s": r["setup_s"],
                "ms_per_step_note": "ms_per_step is the sample's step time scaled to the workload's DoF count"},
        "cpu_baseline": {"value": value, "unit": "DoF-steps/s", "cores": r["threads"], "kind": "port",
                         "sample": sample_txt},
        "e2e": {"value": value, "unit": "DoF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def timed_steps(g, stream, K, flush, torch):
    """K steps of solver g, each bracketed by CUDA events on the context's stream; L2 flushed between."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    its_total = 0
    for k in range(K):
        if flush is not None:
            flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        its, _ = g.step()
        ev[k][1].record(stream)
        its_total += its[0] + its[1]
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev], its_total


def extra_workload(name, stream, flush, torch, K, peak, precond=None):
    """One more BASELINE workload on one GPU: DoF-steps/s, iterations, CG iteration time, flushed SpMV."""
    from wavegpu import WaveSolver, api

    params, scheme, _ = make_params(name, 1)
    t0 = time.time()
    g = WaveSolver(params, scheme, stream=stream.cuda_stream, cg=dict(precond=2) if precond == "mg" else None)
    g.init()
    setup_s = time.time() - t0
    for _ in range(3):
        g.step()
    g.cg_stats(reset=True)
    ms, its = timed_steps(g, stream, K, flush, torch)
    cgs = g.cg_stats()
    info = g.operator_info()
    ms_sp, by_sp = g.bench_spmv(api.MAT_SYS1, reps=10, flush_l2=True)
    out = {"workload": name, "n_dofs": g.n, "value": g.n * K / (sum(ms) * 1e-3), "unit": "DoF-steps/s",
           "ms_per_step": sum(ms) / K, "steps": K, "cg_its_per_step": its / K,
           "cg_ms_per_iteration": cgs["ms_total"] / max(cgs["iterations"], 1),
           "cg_path": "fused (K6f)" if g.cg_fused_active() else "three-kernel",
           "operator": info, "setup_s": setup_s,
           "spmv_l2_flushed": {"ms": ms_sp, "GB/s": by_sp / ms_sp / 1e6, "frac": by_sp / ms_sp / 1e6 / peak,
                               "algorithmic_bytes": by_sp}}
    if precond == "mg":
        out["preconditioner"] = "geometric multigrid V(2,2), damped Jacobi smoothing (WAVE_PRECOND_MG)"
    g.close()
    return out


def measure_e2e(g, api, torch, dist, world, scheme, n, K, barrier):
    """The stateless host-buffer entry point: H2D of u, v, a from pinned memory + step + D2H every step."""
    def pinned(vec):
        t = torch.empty(n, dtype=torch.float64).pin_memory()
        t.numpy()[g.row0:g.row0 + g.nown] = vec
        return t.numpy()

    # every rank keeps (global-length) host arrays and moves its own rows each step
    if world == 1:
        u0_, v0_ = g.vector(api.VEC_U), g.vector(api.VEC_V)
        a0_ = g.vector(api.VEC_A) if scheme == "newmark" else None
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
    return {"value": n * Ke / e2e_s, "unit": "DoF-steps/s", "h2d_bytes_per_step": nvec * 8 * n,
            "d2h_bytes_per_step": nvec * 8 * n + 16 * world, "steps": Ke,
            "api": "wave_step_host (pinned host u, v, a in; u, v, a, norms out; every rank moves its own rows)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precond", default="jacobi", choices=["jacobi", "mg"],
                    help="CG preconditioner: jacobi (north-star default) or the multigrid V-cycle")
    ap.add_argument("--cg", default="auto", choices=["auto", "three-kernel", "fused"],
                    help="CG iteration: auto (the cooperative kernel K6f where the rows fit on chip, else three "
                         "kernels), or force one of them")
    ap.add_argument("--no-stencil", action="store_true", help="keep every matrix row in SELL form")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-extras", "--no-scale-probe", dest="no_extras", action="store_true",
                    help="skip the extra BASELINE workloads (c2, c3, c4, multigrid) reported at N=1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.cg != "auto":
        os.environ["WAVE_CG_FUSED"] = "1" if args.cg == "fused" else "0"  # read by wave_setup
    if args.no_stencil:
        os.environ["WAVE_NO_STENCIL"] = "1"

    import numpy as np  # noqa: F401
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
    info = g.operator_info()
    n = g.n
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up through the timed and the bracketed pass
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        g.step()
    # ---- timed region: K steps, device timed on the context stream -------------------------------
    g.cg_stats(reset=True)
    launches0 = g.launch_count()
    barrier()
    wall0 = time.time()
    step_ms, its_total = timed_steps(g, stream, K, flush, torch)
    barrier()
    wall_s = time.time() - wall0
    total_ms = float(sum(step_ms))
    launches = g.launch_count() - launches0
    cgs = g.cg_stats()
    # second pass over further steps of the same run with every kernel of the CG iterations bracketed by
    # events on the context stream (kept out of the pass above: the brackets cost ~1 us per launch)
    g.kernel_timing(True)
    Kr = max(3, min(K, 10))
    ms2, _ = timed_steps(g, stream, Kr, flush, torch)
    barrier()
    kt_n, kt_ms = g.kernel_timing(False)
    bracketed_ms = float(sum(ms2))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = n * K / (total_ms * 1e-3)

    # ---- e2e: the stateless host-buffer entry point, H2D + D2H of the state every step -------------
    e2e = None
    try:
        e2e = measure_e2e(g, api, torch, dist if world > 1 else None, world, scheme, n, K, barrier)
    except Exception as ex:  # noqa: BLE001
        if world > 1:
            raise  # the other ranks are inside the same collective calls
        torch.cuda.synchronize()
        e2e = {"value": None, "unit": "DoF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "error": f"{type(ex).__name__}: {ex}"}

    # ---- roofline of the dominant kernel, timed live inside the steps above --------------------------
    peak, peak_src = hbm_peak()
    nnz = g.nnz_local
    kernels = []
    # kernel names as the ncu launch lists under profiles/ show them
    names = (("k_spmv_st" if info["stencil_rows"] else "k_spmv") + " (CG A*d with fused d.Ad)",
             "k_cg_update (g += alpha Ad, h = D^-1 g, g.g, g.h)", "k_cg_direction (x += alpha d, d = beta d - h)")
    alg = (float(info["spmv_bytes"]), 40.0 * g.nown, 40.0 * g.nown)
    for k in range(3):
        if kt_n[k] > 0:
            avg = kt_ms[k] / kt_n[k]
            kernels.append({"kernel": names[k], "launches_timed": kt_n[k], "avg_launch_ms_in_step": avg,
                            "algorithmic_bytes_per_launch": alg[k], "achieved": alg[k] / (avg * 1e-3) / 1e9,
                            "frac": alg[k] / (avg * 1e-3) / 1e9 / peak,
                            "share_of_step_time": kt_ms[k] / bracketed_ms if bracketed_ms else None})
    ms_fl, by_fl = g.bench_spmv(api.MAT_SYS1, reps=20, flush_l2=True)
    ms_it, by_it = (float("nan"), float("nan")) if cg_fused else g.bench_cg_iter(api.MAT_SYS1, reps=2)
    traffic = None
    tf = ROOT / "profiles" / "spmv_traffic.json"
    roofline = None
    if kernels:
        dom = max(kernels, key=lambda kk: kk["share_of_step_time"] or 0.0)
        if tf.exists() and world == 1:  # dram bytes per launch from the committed ncu capture of this workload
            traffic = json.loads(tf.read_text()).get(workload, {}).get(dom["kernel"].split(" ")[0])
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak,
                    "peak_source": peak_src, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic,
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                    "avg_launch_ms_in_step": dom["avg_launch_ms_in_step"], "launches_timed": dom["launches_timed"],
                    "share_of_step_time": dom["share_of_step_time"],
                    "timed_over": f"{Kr} further steps of the same run, every CG kernel launch bracketed by CUDA events",
                    "kernels": kernels}
    elif cg_fused:
        # K6f: no per-kernel launches to bracket; the figure is the whole solve (start residual, every
        # iteration, the host's read-back) over its iterations, against its bytes per iteration
        fb = float(info["spmv_bytes"]) - 16.0 * g.nown + 26.0 * g.nown
        ms_i = cgs["ms_total"] / cgs["iterations"] if cgs["iterations"] else float("nan")
        roofline = {"bound": "hbm", "kernel": "k_cg_fused (one cooperative kernel per solve: SpMV from staged d, "
                                              "updates on chip)",
                    "achieved": fb / (ms_i * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": fb / (ms_i * 1e-3) / 1e9 / peak, "traffic": None,
                    "algorithmic_bytes_per_launch": fb, "avg_launch_ms_in_step": ms_i,
                    "launches_timed": cgs["iterations"],
                    "share_of_step_time": cgs["ms_total"] / total_ms if total_ms else None,
                    "timed_over": "all CG solves of the timed steps: device ms per solve / iterations"}
    if roofline is not None and info["stencil_rows"] and kernels:
        # the stencil SpMV computes the product a CSR kernel would stream 12 nnz + 20 n bytes for
        csr_bytes = 12.0 * nnz + 20.0 * g.nown
        sp = kernels[0]
        roofline["spmv_csr_equivalent"] = {
            "bytes": csr_bytes, "GB/s": csr_bytes / (sp["avg_launch_ms_in_step"] * 1e-3) / 1e9,
            "x_peak": csr_bytes / (sp["avg_launch_ms_in_step"] * 1e-3) / 1e9 / peak,
            "note": "k_spmv_st's frac is against the 16 B/row the table-driven operator still has to move "
                    "(x in, y out); run with --no-stencil for the SELL kernel's own roofline"}
    if roofline is not None:
        roofline["spmv_l2_flushed_single_launch"] = {"ms": ms_fl, "GB/s": by_fl / ms_fl / 1e6,
                                                     "frac": by_fl / ms_fl / 1e6 / peak, "algorithmic_bytes": by_fl}
        roofline["cg_iteration"] = None if cg_fused else {"ms": ms_it, "algorithmic_bytes": by_it,
                                                         "GB/s": by_it / ms_it / 1e6,
                                                         "frac": by_it / ms_it / 1e6 / peak}
        roofline["operator"] = info
    g.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the other BASELINE workloads on one GPU, and the multigrid preconditioner ---------------------
    extras = None
    if world == 1 and not args.no_extras:
        Kx = max(3, min(K, 10))
        extras = []
        # the headline line must not depend on an extra: a failure there is reported in its place
        def safely(name, **kw):
            try:
                return extra_workload(name, stream, flush, torch, Kx, peak, **kw)
            except Exception as ex:  # noqa: BLE001
                torch.cuda.synchronize()
                return {"workload": name, "error": f"{type(ex).__name__}: {ex}", **kw}

        for name in EXTRAS:
            if name != workload:
                extras.append(safely(name))
        if args.precond == "jacobi":
            extras.append(safely(workload, precond="mg"))
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            sample, what = cpu_sample_params(params)
            r = time_oracle(sample, scheme, budget_s=25.0, max_steps=10, precond=2 if args.precond == "mg" else 0)
            cpu = {"value": r["n"] * r["steps"] / r["seconds"], "unit": "DoF-steps/s", "cores": r["threads"],
                   "kind": "port", "cg_its_per_step": r["cg_its_per_step"],
                   "sample": f"{what}: {r['n']} DoFs, {r['steps']} time steps on the host "
                             f"(oracle/wave_oracle.c, OpenMP, {r['threads']} threads)"}
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": "DoF-steps/s", "cores": 0, "kind": "port",
                   "sample": f"CPU leg failed: {type(ex).__name__}: {ex}"}

    line = {
        "metric": "dof_steps_per_sec", "value": value, "unit": "DoF-steps/s", "n_gpus": n_gpus, "steps": K,
        "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": (value / PUBLISHED[workload]) if workload in PUBLISHED else None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(workload, params, scheme, n, args.precond),
        "run": {"nnz_per_gpu": nnz, "cg_its_per_step": its_total / K,
                "cg_path": "fused (K6f)" if cg_fused else "three-kernel", "parallelism": f"strips{n_gpus}",
                "operator": "stencil tables + SELL-32" if info["stencil_rows"] else "SELL-32",
                "l2": "flushed between steps (256 MiB write)" if flush is not None else "not flushed",
                "setup_s": setup_s, "wall_s_timed_region": wall_s},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        "extras": extras, "cpu_baseline": cpu,
        "cg": {"solves": cgs["solves"], "iterations": cgs["iterations"], "ms_total": cgs["ms_total"],
               "ms_per_iteration": cgs["ms_total"] / max(cgs["iterations"], 1)},
        "step_ms": {"min": min(step_ms), "max": max(step_ms)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
