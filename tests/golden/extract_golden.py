"""Extract the known-answer rows this repo tests against from the reference's own shipped
result tables (run once in the build container; /root/reference does not exist on the GPU box).

    python tests/golden/extract_golden.py [--all]

Sources (reference artefacts, read-only):
  analysis/data/convergence-results.csv   final relative L2/H1 errors, standing mode, T=1
      recipe: scripts/convergence_sweep.py:165-179 (standing-mode-wsol.json + overrides)
  analysis/data/dissdisp-results.csv      energy_ratio etc., Nel=60, r=1, T=5, Log Every=1
      recipe: scripts/dissipation_dispersion_sweep.py:179-198
Outputs: tests/golden/convergence_rows.json, tests/golden/dissdisp_rows.json (the rows the test suite
runs) and, with --all, convergence_rows_all.json / dissdisp_rows_all.json (every row, for the offline sweeps)
and scalability_seconds.json (mean wall seconds per scheme and rank count of analysis/data/scalability-results.csv)
"""
import csv
import json
from pathlib import Path

REF = Path("/root/reference/analysis/data")
OUT = Path(__file__).resolve().parent


def main():
    rows = []
    with open(REF / "convergence-results.csv") as fh:
        for lineno, row in enumerate(csv.DictReader(fh), start=2):
            nel, r, dt = int(row["N_el_x"]), int(row["r"]), float(row["dt"])
            l2, h1 = float(row["rel_L2_error_final"]), float(row["rel_H1_error_final"])
            # small, cheap and stable cases (unstable explicit runs amplify round-off: not a known answer)
            if nel > 40 or dt < 0.001 or l2 > 1.5 or (nel == 40 and (dt < 0.005 or r == 2)):
                continue
            # explicit schemes (theta=0 / beta=0) above half the CFL bound of scripts/convergence_sweep.py:139-147
            # grow round-off noise seeded by the reference's 16-rank numbering: not a known answer
            explicit = row["theta"] == "0.000000" or row["beta"] == "0.000000"
            cfl = 0.9 / nel / (2 ** 0.5 * (1.0 if r == 1 else 4.0))
            if explicit and dt > 0.5 * cfl:
                continue
            method = "theta" if row["method"].startswith("theta") else "newmark"
            rows.append({
                "line": lineno, "scheme": method, "Nel": nel, "R": r, "Dt": row["dt"], "T": row["T"],
                "Theta": None if row["theta"] == "N/A" else float(row["theta"]),
                "Beta": None if row["beta"] == "N/A" else float(row["beta"]),
                "Gamma": None if row["gamma"] == "N/A" else float(row["gamma"]),
                "rel_L2": l2, "rel_H1": h1})
    (OUT / "convergence_rows.json").write_text(json.dumps(rows, indent=0))
    drows = []
    with open(REF / "dissdisp-results.csv") as fh:
        for lineno, row in enumerate(csv.DictReader(fh), start=2):
            if float(row["dt"]) < 0.01 or float(row["max_rel_L2"]) > 50:
                continue
            drows.append({"line": lineno, "scheme": row["scheme"], "Nel": int(row["Nel"]), "R": int(row["R"]),
                          "Dt": row["dt"], "T": row["T"], "energy_ratio": float(row["energy_ratio"]),
                          "max_rel_L2": float(row["max_rel_L2"]), "final_rel_L2": float(row["final_rel_L2"]),
                          "final_rel_H1": float(row["final_rel_H1"])})
    (OUT / "dissdisp_rows.json").write_text(json.dumps(drows, indent=0))
    print(len(rows), "convergence rows,", len(drows), "dissdisp rows")


def main_all():
    """Every row of both tables, unfiltered (tools/golden_sweep_oracle.py, tools/golden_sweep_gpu.py --all)."""
    rows = []
    with open(REF / "convergence-results.csv") as fh:
        for lineno, row in enumerate(csv.DictReader(fh), start=2):
            method = "theta" if row["method"].startswith("theta") else "newmark"
            rows.append({
                "line": lineno, "scheme": method, "Nel": int(row["N_el_x"]), "R": int(row["r"]), "Dt": row["dt"],
                "T": row["T"],
                "Theta": None if row["theta"] == "N/A" else float(row["theta"]),
                "Beta": None if row["beta"] == "N/A" else float(row["beta"]),
                "Gamma": None if row["gamma"] == "N/A" else float(row["gamma"]),
                "rel_L2": float(row["rel_L2_error_final"]), "rel_H1": float(row["rel_H1_error_final"])})
    (OUT / "convergence_rows_all.json").write_text(json.dumps(rows, indent=0))
    drows = []
    with open(REF / "dissdisp-results.csv") as fh:
        for lineno, row in enumerate(csv.DictReader(fh), start=2):
            drows.append({"line": lineno, "scheme": row["scheme"], "Nel": int(row["Nel"]), "R": int(row["R"]),
                          "Dt": row["dt"], "T": row["T"], "energy_ratio": float(row["energy_ratio"]),
                          "max_rel_L2": float(row["max_rel_L2"]), "final_rel_L2": float(row["final_rel_L2"]),
                          "final_rel_H1": float(row["final_rel_H1"])})
    (OUT / "dissdisp_rows_all.json").write_text(json.dumps(drows, indent=0))
    print(len(rows), "convergence rows,", len(drows), "dissdisp rows (all)")
    # analysis/data/scalability-results.csv (scripts/scalability_sweep.py:183-230): whole-process wall seconds,
    # Nel=640, R=1, Dt=8e-5, T=0.05; mean over the repeats per (scheme, nprocs)
    acc = {}
    with open(REF / "scalability-results.csv") as fh:
        for row in csv.DictReader(fh):
            if row["returncode"] == "0":
                acc.setdefault((row["scheme"], int(row["nprocs"])), []).append(float(row["seconds"]))
    table = {}
    for (scheme, nprocs), secs in sorted(acc.items()):
        table.setdefault(scheme, {})[str(nprocs)] = round(sum(secs) / len(secs), 3)
    (OUT / "scalability_seconds.json").write_text(json.dumps(table, indent=0))
    print(len(acc), "scalability (scheme, nprocs) means")


if __name__ == "__main__":
    import sys

    main()
    if "--all" in sys.argv:
        main_all()
