"""Kernel-level probe used for ncu captures and roofline checks:
    python tools/spmv_probe.py --nel 2048 --r 2 [--reps 20] [--cg]
Prints one JSON line with the SpMV launch time (L2 flushed / back to back) and, with --cg, the
average device time of a Jacobi-PCG iteration."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nel", default="1024")
    ap.add_argument("--r", type=int, default=1)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--scheme", default="newmark")
    ap.add_argument("--dt", default="0.01")
    ap.add_argument("--cg", action="store_true")
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--no-init", action="store_true", help="skip wave_init: only the SpMV launches (ncu captures)")
    args = ap.parse_args()
    from wavegpu import WaveSolver, api, problem

    p = problem("standing-mode-wsol", Nel=args.nel, R=args.r, Dt=args.dt)
    g = WaveSolver(p, args.scheme)
    if not args.no_init:
        g.init()
    peak = 6549.4
    out = {"nel": args.nel, "r": args.r, "n": g.n, "nnz": g.nnz_local}
    ms, nbytes = g.bench_spmv(api.MAT_SYS1, reps=args.reps, flush_l2=False)
    out["spmv_back_to_back"] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    ms, nbytes = g.bench_spmv(api.MAT_SYS1, reps=min(args.reps, 10), flush_l2=True)
    out["spmv_l2_flushed"] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    if args.cg:
        ms, nbytes = g.bench_cg_iter(api.MAT_SYS1, reps=2)
        out["cg_iteration"] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    if args.steps:
        g.timers_enable(True)
        done, its, nrm, tot = g.run(args.steps)
        out["steps"] = {"n": done, "cg_its": tot, "timers_ms": g.timers()}
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()
