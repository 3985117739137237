"""Offline sweep of the CPU oracle over EVERY row of the reference's result tables
(tests/golden/*_all.json: 490 rows of analysis/data/convergence-results.csv, 47 rows of
dissdisp-results.csv), cheapest rows first, in parallel worker processes, until a time budget runs out.
The test suite runs the 136 cheap, stable rows; this tool documents how far the agreement extends.

    python tools/golden_sweep_oracle.py --minutes 120 --workers 7 --out profiles/r1_oracle_full_sweep.json
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor, as_completed
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))
TIGHT = dict(reduce=1e-13, tol=1e-30)


def n_steps(dt, T):
    t, n = 0.0, 0
    while t < T:
        t += dt
        n += 1
    return n


def cost(row):
    nel, r = row["Nel"], row["R"]
    n = (r * nel + 1) ** 2
    return n_steps(float(row["Dt"]), float(row["T"])) * n * (1.0 if r == 1 else 1.7)


def run_conv(row):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle as O
    from wavegpu.problems import problem

    kw = dict(Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"])
    for k in ("Theta", "Beta", "Gamma"):
        if row.get(k) is not None:
            kw[k] = row[k]
    t0 = time.time()
    out = O.run(problem("standing-mode-wsol", **kw), row["scheme"], cg=TIGHT)
    _, _, l2, h1 = out["final_errors"]
    return {"table": "convergence", "line": row["line"], "scheme": row["scheme"], "Nel": row["Nel"], "R": row["R"],
            "Dt": row["Dt"], "Theta": row["Theta"], "Beta": row["Beta"], "gold": [row["rel_L2"], row["rel_H1"]],
            "oracle": [l2, h1], "dev": [abs(l2 - row["rel_L2"]) / row["rel_L2"], abs(h1 - row["rel_H1"]) / row["rel_H1"]],
            "diverged": out.get("diverged"), "seconds": time.time() - t0}


def run_diss(row):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle as O
    from wavegpu.problems import problem

    kind, val = row["scheme"].split("-")
    scheme = "theta" if kind == "theta" else "newmark"
    extra = {"Theta": val} if kind == "theta" else {"Beta": val, "Gamma": "0.5"}
    p = problem("standing-mode-wsol", Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"], **extra)
    t0 = time.time()
    out = O.run(p, scheme, log_every=1, cg=TIGHT)
    E = [float("%.6g" % e[2]) for e in out["energy"]]
    rl2 = [e[4] for e in out["error"]]
    got = [E[-1] / E[0], max(rl2), rl2[-1], out["error"][-1][5]]
    gold = [row["energy_ratio"], row["max_rel_L2"], row["final_rel_L2"], row["final_rel_H1"]]
    return {"table": "dissdisp", "line": row["line"], "scheme": row["scheme"], "Dt": row["Dt"], "gold": gold,
            "oracle": got, "dev": [abs(a - b) / abs(b) if b else abs(a) for a, b in zip(got, gold)],
            "diverged": out.get("diverged"), "seconds": time.time() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=60.0)
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--out", default=str(ROOT / "profiles" / "oracle_full_sweep.json"))
    args = ap.parse_args()
    gold = ROOT / "tests" / "golden"
    conv = json.loads((gold / "convergence_rows_all.json").read_text())
    diss = json.loads((gold / "dissdisp_rows_all.json").read_text())
    jobs = [(cost(r), run_conv, r) for r in conv]
    jobs += [(n_steps(float(r["Dt"]), float(r["T"])) * (r["Nel"] + 1) ** 2 * 3.0, run_diss, r) for r in diss]
    jobs.sort(key=lambda j: j[0])
    deadline = time.time() + 60.0 * args.minutes
    results, skipped = [], 0
    with ProcessPoolExecutor(max_workers=args.workers) as ex:
        pending = {}
        it = iter(jobs)

        def submit_next():
            for c, fn, row in it:
                pending[ex.submit(fn, row)] = c
                return True
            return False

        for _ in range(args.workers):
            submit_next()
        while pending:
            for fut in as_completed(list(pending)):
                pending.pop(fut)
                results.append(fut.result())
                if time.time() < deadline:
                    submit_next()
                break
            Path(args.out + ".partial").write_text(json.dumps(results))
        skipped = sum(1 for _ in it)
    results.sort(key=lambda r: (r["table"], r["line"]))
    summary = {"command": " ".join(sys.argv), "rows_run": len(results), "rows_not_reached": skipped,
               "cg": "tight (reduce 1e-13)", "rows": results}
    Path(args.out).write_text(json.dumps(summary, indent=0))
    Path(args.out + ".partial").unlink(missing_ok=True)
    print(len(results), "rows run,", skipped, "not reached")


if __name__ == "__main__":
    main()
