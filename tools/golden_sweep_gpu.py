"""Run every known-answer row of tests/golden/ through libwavegpu (the reference's own result tables:
analysis/data/convergence-results.csv and dissdisp-results.csv) and report the worst deviations.
    python tools/golden_sweep_gpu.py            # on a B200: the 136 rows of the test suite
    python tools/golden_sweep_gpu.py --all OUT  # every row of both tables (490 + 47), per-row JSON to OUT
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))
from wavegpu import WaveSolver, problem  # noqa: E402

TIGHT = dict(reduce=1e-13, tol=1e-30)


def run(p, scheme, log_every=0):
    g = WaveSolver(p, scheme, cg=TIGHT)
    try:
        g.init()
        t, dt, T, step = 0.0, float(p["Dt"]), float(p["T"]), 0
        energies, rel_l2 = [], []
        while t < T:
            t += dt
            step += 1
            g.step()
            if log_every and step % log_every == 0:
                energies.append(float("%.6g" % g.energy()))
                rel_l2.append(g.errors()[2])
        return g.errors(), energies, rel_l2
    finally:
        g.close()


def main():
    t0 = time.time()
    conv = json.loads((ROOT / "tests" / "golden" / "convergence_rows.json").read_text())
    worst = {1: [0.0, 0.0], 2: [0.0, 0.0]}
    for row in conv:
        kw = dict(Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"])
        for k in ("Theta", "Beta", "Gamma"):
            if row.get(k) is not None:
                kw[k] = row[k]
        e, _, _ = run(problem("standing-mode-wsol", **kw), row["scheme"])
        d2, d1 = abs(e[2] - row["rel_L2"]) / row["rel_L2"], abs(e[3] - row["rel_H1"]) / row["rel_H1"]
        worst[row["R"]][0] = max(worst[row["R"]][0], d2)
        worst[row["R"]][1] = max(worst[row["R"]][1], d1)
    diss = json.loads((ROOT / "tests" / "golden" / "dissdisp_rows.json").read_text())
    worst_ratio, exact_ratio = 0.0, 0
    for row in diss:
        kind, val = row["scheme"].split("-")
        extra = {"Theta": val} if kind == "theta" else {"Beta": val, "Gamma": "0.5"}
        p = problem("standing-mode-wsol", Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"], **extra)
        e, E, rl2 = run(p, "theta" if kind == "theta" else "newmark", log_every=1)
        ratio = E[-1] / E[0]
        exact_ratio += int(ratio == row["energy_ratio"])
        worst_ratio = max(worst_ratio, abs(ratio - row["energy_ratio"]) / row["energy_ratio"])
    print(json.dumps({"convergence_rows": len(conv),
                      "worst_rel_dev_r1": {"rel_L2": worst[1][0], "rel_H1": worst[1][1]},
                      "worst_rel_dev_r2": {"rel_L2": worst[2][0], "rel_H1": worst[2][1]},
                      "dissdisp_rows": len(diss), "energy_ratio_bit_identical_rows": exact_ratio,
                      "worst_energy_ratio_dev": worst_ratio, "seconds": time.time() - t0}, indent=1))


def main_all(out_path):
    """Every row of both tables (the reference's scripts/*_sweep.py regression, from the fixtures):
    per-row oracle-free comparison with the printed values; long rows take minutes on a B200."""
    gold = ROOT / "tests" / "golden"
    rows = []
    for row in json.loads((gold / "convergence_rows_all.json").read_text()):
        kw = dict(Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"])
        for k in ("Theta", "Beta", "Gamma"):
            if row.get(k) is not None:
                kw[k] = row[k]
        t0 = time.time()
        try:
            e, _, _ = run(problem("standing-mode-wsol", **kw), row["scheme"])
            got, err = [e[2], e[3]], None
        except Exception as ex:  # diverged explicit runs end in WAVE_ERR_DIVERGED
            got, err = None, str(ex)
        rows.append({"table": "convergence", "line": row["line"], "gold": [row["rel_L2"], row["rel_H1"]], "gpu": got,
                     "error": err, "seconds": time.time() - t0})
        Path(out_path).write_text(json.dumps(rows))
    for row in json.loads((gold / "dissdisp_rows_all.json").read_text()):
        kind, val = row["scheme"].split("-")
        extra = {"Theta": val} if kind == "theta" else {"Beta": val, "Gamma": "0.5"}
        p = problem("standing-mode-wsol", Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"], **extra)
        t0 = time.time()
        try:
            e, E, rl2 = run(p, "theta" if kind == "theta" else "newmark", log_every=1)
            got, err = [E[-1] / E[0], max(rl2), rl2[-1], e[3]], None
        except Exception as ex:
            got, err = None, str(ex)
        rows.append({"table": "dissdisp", "line": row["line"],
                     "gold": [row["energy_ratio"], row["max_rel_L2"], row["final_rel_L2"], row["final_rel_H1"]],
                     "gpu": got, "error": err, "seconds": time.time() - t0})
        Path(out_path).write_text(json.dumps(rows))
    print(len(rows), "rows ->", out_path)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--all":
        main_all(sys.argv[2])
    else:
        main()
