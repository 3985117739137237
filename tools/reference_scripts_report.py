#!/usr/bin/env python
"""Run the reference's own sweep drivers UNMODIFIED against this repository's launcher and host executables,
and compare the tables they write with the reference's result tables.

The drivers (scripts/convergence_sweep.py:86-96, scripts/dissipation_dispersion_sweep.py:113-122,
scripts/scalability_sweep.py:71-85) locate everything relative to their own file: `../parameters/
standing-mode-wsol.json`, `../build/main-theta`, `../build/main-newmark`, and start
`<launcher> -np P [--bind-to core --map-by socket] <binary> <params.json>` from `build/`.  `checkout()` lays
that directory tree out in a scratch directory -- the scripts and parameter files are copied there from the
reference at run time, never into this repository -- with `build/main-*` pointing at the executables to test,
and `run_script()` starts a driver with `--launcher tools/mpirun-shim` (the product's wave-mpirun).

Where there is a GPU, `build/main-*` are the product's executables (nmpde-wave-equation_b200/bin/).  The
reference does not exist on the GPU box, so in this container -- no GPU -- the drivers run against the host
classes linked with the TEST DOUBLE of the C ABI (tests/abi_double/, numerics = the CPU oracle): everything
above the ABI (launcher, parameter reader, folder and CSV conventions, exit codes) is the product's code.

    python tools/reference_scripts_report.py run  --reference /root/reference --work /tmp/sweeps \\
        --script convergence_sweep.py -- --nprocs 4 --nel 10 20 40
    python tools/reference_scripts_report.py report --conv a.csv [b.csv ...] --diss d.csv --scal s.csv \\
        --out profiles/r2_reference_scripts.md
"""
import argparse
import csv
import json
import math
import os
import shutil
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "tools" / "mpirun-shim"
TIGHT_CG = {"WAVE_CG_REDUCE": "1e-13", "WAVE_CG_TOL": "1e-30"}  # the reference's AMG-CG lands far below its 1e-6 bar
BINS = [1e-6, 1e-5, 1e-4, 1e-3]


def checkout(reference, work, bin_dir):
    """The directory tree the drivers expect, in `work`: scripts/ and parameters/ copied from the reference,
    build/main-theta and build/main-newmark -> the executables in `bin_dir`."""
    reference, work, bin_dir = Path(reference).resolve(), Path(work).resolve(), Path(bin_dir).resolve()
    work.mkdir(parents=True, exist_ok=True)
    for sub in ("scripts", "parameters"):
        if not (work / sub).exists():
            shutil.copytree(reference / sub, work / sub)
    (work / "build").mkdir(exist_ok=True)
    for exe in ("main-theta", "main-newmark"):
        link = work / "build" / exe
        if link.is_symlink() or link.exists():
            link.unlink()
        link.symlink_to(bin_dir / exe)
    return work


def run_script(work, script, script_args, env=None, timeout=None):
    """Start one driver the way its docstring says (from build/), with this repository's launcher."""
    work = Path(work)
    cmd = [sys.executable, str(work / "scripts" / script), "--launcher", str(SHIM), *script_args]
    e = dict(os.environ)
    e.update(env or {})
    t0 = time.time()
    r = subprocess.run(cmd, cwd=work / "build", capture_output=True, text=True, env=e, timeout=timeout)
    return r, time.time() - t0


# ---- comparison with the reference's tables (fixtures tests/golden/*_rows_all.json) ---------------------
def _f(x):
    try:
        return float(x)
    except (TypeError, ValueError):
        return None


def _dev(got, gold):
    worst = 0.0
    for a, b in zip(got, gold):
        if b is None or not math.isfinite(b):
            return None  # the reference's own run blew up: nothing to compare
        if a is None or not math.isfinite(a):
            return float("inf")
        worst = max(worst, abs(a - b) / abs(b) if b != 0 else abs(a))
    return worst


def _class(explicit, gold):
    if any(g is None or not math.isfinite(g) for g in gold):
        return "reference value nan/inf (its run blew up)"
    if explicit:
        return "explicit, printed error > 1.5 (unstable run)" if max(gold[:2]) > 1.5 else \
            "explicit (theta=0, beta=0), printed error <= 1.5"
    return "implicit (theta=1/2, theta=1, beta=1/4)"


def _table(rows):
    classes = {}
    for klass, d in rows:
        c = classes.setdefault(klass, [0] * (len(BINS) + 2))
        c[0] += 1
        if d is None:
            c[-1] += 1
            continue
        for k, b in enumerate(BINS):
            if d <= b:
                c[1 + k] += 1
                break
        else:
            c[-1] += 1
    out = ["| class | rows | <= 1e-6 | <= 1e-5 | <= 1e-4 | <= 1e-3 | worse / not comparable |", "|---|---|---|---|---|---|---|"]
    for name in sorted(classes):
        c = classes[name]
        out.append(f"| {name} | {c[0]} | " + " | ".join(str(x) for x in c[1:]) + " |")
    return out


def compare_convergence(files):
    """Rows of the drivers' merged convergence-results.csv against analysis/data/convergence-results.csv."""
    gold = {}
    for g in json.loads((ROOT / "tests" / "golden" / "convergence_rows_all.json").read_text()):
        par = g["Theta"] if g["scheme"] == "theta" else g["Beta"]
        gold[(g["scheme"], g["Nel"], g["R"], float(g["Dt"]), float(par))] = g
    rows, unmatched, seen = [], 0, set()
    for f in files:
        for r in csv.DictReader(Path(f).open()):
            scheme = "theta" if r["method"].startswith("theta") else "newmark"
            par = float(r["theta"]) if scheme == "theta" else float(r["beta"])
            key = (scheme, int(r["N_el_x"]), int(r["r"]), float(r["dt"]), par)
            g = gold.get(key)
            if g is None or key in seen:
                unmatched += 1
                continue
            seen.add(key)
            want = [g["rel_L2"], g["rel_H1"]]
            d = _dev([_f(r["rel_L2_error_final"]), _f(r["rel_H1_error_final"])], want)
            rows.append({"key": key, "line": g["line"], "class": _class(par == 0.0, want), "dev": d})
    return rows, unmatched, len(gold)


def compare_dissdisp(files):
    files = [files] if isinstance(files, (str, Path)) else list(files)
    gold = {}
    for g in json.loads((ROOT / "tests" / "golden" / "dissdisp_rows_all.json").read_text()):
        gold[(g["scheme"], g["Nel"], g["R"], float(g["Dt"]))] = g
    rows, unmatched = [], 0
    for r in (row for f in files for row in csv.DictReader(Path(f).open())):
        key = (r["scheme"], int(r["Nel"]), int(r["R"]), float(r["dt"]))
        g = gold.get(key)
        if g is None:
            unmatched += 1
            continue
        names = ("energy_ratio", "max_rel_L2", "final_rel_L2", "final_rel_H1")
        want = [g[k] for k in names]
        got = [_f(r[k]) for k in names]
        if all(v is None for v in got):  # the driver gave up on the run (its --timeout): no numbers in its table
            unmatched += 1
            continue
        rows.append({"key": key, "line": g["line"], "class": _class(key[0] in ("theta-0.0", "newmark-0.00"), want[1:]),
                     "dev": _dev(got, want), "energy_ratio": (got[0], want[0])})
    return rows, unmatched, len(gold)


def _conv_section(title, files, how, out):
    rows, unmatched, total = compare_convergence(files)
    out += [f"### convergence_sweep.py: {len(rows)} of the {total} rows of analysis/data/convergence-results.csv regenerated", "",
            f"Arguments: `{how}`.  Deviation = |regenerated - printed| / printed, the larger of the final relative L2 and",
            "H1 errors (7 printed digits).", "", *_table([(r["class"], r["dev"]) for r in rows]), ""]
    fin = [r for r in rows if r["dev"] is not None and math.isfinite(r["dev"])]
    exact = sum(1 for r in fin if r["dev"] <= 5e-7)
    out += [f"{exact} rows reproduce every printed digit (deviation <= 5e-7, the rounding of a 7-digit number).  "
            f"Rows the driver wrote that the table does not hold: {unmatched}.  (Classes as in `profiles/r2_golden_all.md`:",
            "explicit runs above their stability limit amplify round-off without bound, so their printed errors -- and the",
            "reference's nan/inf rows -- are not known answers.)", ""]
    worst = sorted((r for r in fin if r["class"].startswith("implicit") and r["dev"] > 1e-6), key=lambda r: -r["dev"])[:5]
    if worst:
        out += ["Largest deviations among the implicit rows: P2 rows with printed errors around 1e-6, where the reference's own",
                "theta = 1/2 and Newmark-1/4 rows -- mathematically the same scheme for F = 0, g = 0 -- differ by as much because",
                "its CG stops at a 1e-6 residual reduction (e.g. csv line 147, theta = 1/2: 1.139029e-06 against line 451,",
                "Newmark 1/4: 1.104551e-06, same Nel = 80, R = 2, dt = 1e-4; regenerated here: 1.104552e-06 for both):", "",
                "| csv line | scheme | Nel | R | dt | theta / beta | deviation |", "|---|---|---|---|---|---|---|"]
        out += [f"| {r['line']} | {r['key'][0]} | {r['key'][1]} | {r['key'][2]} | {r['key'][3]:g} | {r['key'][4]:g} | {r['dev']:.1e} |"
                for r in worst] + [""]
    return rows


def _diss_section(f, how, out):
    rows, unmatched, total = compare_dissdisp(f)
    out += [f"### dissipation_dispersion_sweep.py: {len(rows)} of the {total} rows of analysis/data/dissdisp-results.csv regenerated", "",
            f"Arguments: `{how}` (Nel = 60, R = 1, T = 5, logging every step; the driver computes the energy ratio and the",
            "maximal / final errors from the `energy.csv` / `error.csv` / `probe.csv` it finds under the run folder it predicts).",
            "Deviation = the largest over energy ratio, max / final rel-L2, final rel-H1.", "",
            *_table([(r["class"], r["dev"]) for r in rows]), ""]
    same = sum(1 for r in rows if r["energy_ratio"][0] == r["energy_ratio"][1])
    out += [f"Energy ratios identical to the table's to the last bit: {same} of {len(rows)}.  "
            f"Rows without a partner in the table or without numbers (run stopped by the driver's own --timeout): {unmatched}.", ""]
    return rows


def report(args):
    out = ["# The reference's sweep drivers, unmodified, against this repository's launcher and executables (row f3)",
           "",
           "`tools/reference_scripts_report.py`: `scripts/convergence_sweep.py`, `scripts/dissipation_dispersion_sweep.py`",
           "and `scripts/scalability_sweep.py`, copied at run time from the reference into a scratch checkout layout",
           "(`scripts/`, `parameters/`, `build/main-theta`, `build/main-newmark`) and started from `build/` with",
           "`--launcher tools/mpirun-shim` -- no edit to any of them; the copies are never committed.  The launcher is the",
           "product's `wave-mpirun` (mpirun's command line: `-np 4 --bind-to core --map-by socket <binary> <file>` as the",
           "drivers build it; one rank per visible GPU).  CG solved tightly (`WAVE_CG_REDUCE=1e-13`): the reference's",
           "AMG-preconditioned solves land far below their 1e-6 stopping bar, Jacobi-CG stops at it.",
           ""]
    if args.b200_conv or args.b200_diss:
        out += ["## 1. On a B200, through the product's executables and libwavegpu", "",
                "`tools/run_reference_drivers.sh nmpde-wave-equation_b200/bin <scratch copy> gpurun_out/…` on one B200",
                f"({args.b200_note}).  Every run is a fresh process (CUDA context + set-up: 2.3-5.8 s of wall time per run as",
                "the drivers log it, for 0.004-0.07 s inside the time loop), so the grid is small; raw files:",
                "`profiles/r2_reference_scripts_b200/`.", ""]
        if args.b200_conv:
            _conv_section("b200", args.b200_conv, args.b200_conv_args, out)
        if args.b200_diss:
            _diss_section(args.b200_diss, args.b200_diss_args, out)
    if args.conv or args.diss or args.scal:
        out += ["## 2. In the build container (no GPU), host classes on the test double of the C ABI", "",
                "The larger grids.  `build/main-*` are the product's host sources (`nmpde-wave-equation_b200/host/*.cpp`,",
                "unchanged: parameter reader, folder naming, CSV writers, exit codes) linked with",
                "`tests/abi_double/wave_abi_on_oracle.cpp` -- the ~25 entry points the host classes call, implemented on the CPU",
                "oracle -- instead of libwavegpu.so (test infrastructure, built into `tests/_build/` only).  This exercises",
                "everything *above* the ABI with the drivers' full parameter ranges; *below* the ABI the same rows through",
                "libwavegpu are section 1 and `profiles/r2_golden_all.md` (all 537 rows, library calls).  The same flow is a",
                "test: `tests/test_reference_scripts_cpu.py`.", ""]
        if args.conv:
            _conv_section("cpu", args.conv, args.conv_args, out)
        if args.diss:
            _diss_section(args.diss, args.diss_args, out)
        if args.scal:
            out += ["### scalability_sweep.py --nprocs 1", "",
                    "The driver's fixed configuration (Nel = 640, R = 1, Dt = 8e-5, T = 0.05: 410 881 DoFs, 625 steps), whole-",
                    "process wall time per scheme as the driver measures it.  Here that is the time of the **CPU oracle** behind",
                    f"the test double ({args.scal_note}): it says nothing about the GPU and is listed to show that the driver's",
                    "table is produced (return codes 0).  The reference's own figures (AMG-CG, Xeon Gold 6238R) are next to it;",
                    "the same configuration through libwavegpu on a B200 is in `profiles/r1_published_config.md`.", "",
                    "| scheme | returncode | seconds (oracle behind the double) | reference table: 1 rank | 16 ranks | 32 ranks |",
                    "|---|---|---|---|---|---|"]
            ref = json.loads((ROOT / "tests" / "golden" / "scalability_seconds.json").read_text())
            for r in csv.DictReader(Path(args.scal).open()):
                g = ref.get(r["scheme"], {})
                out.append(f"| {r['scheme']} | {r['returncode']} | {float(r['seconds']):.1f} | {g.get('1', '–')} | "
                           f"{g.get('16', '–')} | {g.get('32', '–')} |")
            out.append("")
    Path(args.out).write_text("\n".join(out))
    print(f"wrote {args.out}")


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("run")
    r.add_argument("--reference", default="/root/reference")
    r.add_argument("--work", required=True)
    r.add_argument("--script", required=True)
    r.add_argument("--bin", default=None, help="directory with main-theta / main-newmark (default: build the test double)")
    r.add_argument("--tight-cg", action="store_true")
    r.add_argument("script_args", nargs="*")
    p = sub.add_parser("report")
    p.add_argument("--conv", nargs="*", default=[])
    p.add_argument("--conv-args", default="")
    p.add_argument("--diss", nargs="*", default=[])
    p.add_argument("--diss-args", default="--nprocs 4")
    p.add_argument("--scal", default=None)
    p.add_argument("--scal-note", default="OpenMP")
    p.add_argument("--b200-conv", nargs="*", default=[])
    p.add_argument("--b200-conv-args", default="")
    p.add_argument("--b200-diss", nargs="*", default=[])
    p.add_argument("--b200-diss-args", default="")
    p.add_argument("--b200-note", default="")
    p.add_argument("--out", required=True)
    args = ap.parse_args()
    if args.cmd == "report":
        report(args)
        return
    bin_dir = args.bin
    if bin_dir is None:
        sys.path.insert(0, str(ROOT / "tests" / "abi_double"))
        import build_double

        bin_dir = build_double.build()
    checkout(args.reference, args.work, bin_dir)
    res, secs = run_script(args.work, args.script, args.script_args, env=TIGHT_CG if args.tight_cg else None)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    print(f"[{args.script}] exit {res.returncode} after {secs:.1f} s; outputs in {Path(args.work) / 'build'}")
    sys.exit(res.returncode)


if __name__ == "__main__":
    main()
