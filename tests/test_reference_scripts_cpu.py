"""The reference's own sweep drivers, UNMODIFIED, end to end against this repository's launcher and host
executables.

Runs only where the reference checkout exists (the build container; skipped on the GPU box, which has no
/root/reference -- and reference sources are never copied into this repository: the drivers are copied into
pytest's scratch directory at run time).  There is no GPU in that container, so `build/main-theta` and
`build/main-newmark` are the product's host classes linked with the TEST DOUBLE of the C ABI
(tests/abi_double/, numerics = the CPU oracle); the launcher is the product's (tools/mpirun-shim ->
bin/wave-mpirun).  The drivers' tables are compared with the reference's result tables
(analysis/data/*.csv, fixtures tests/golden/*_rows_all.json)."""
import csv
import math
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path(os.environ.get("WAVE_REFERENCE_ROOT", "/root/reference"))
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.skipif(not (REFERENCE / "scripts" / "convergence_sweep.py").exists(),
                                reason="the reference checkout is not on this machine")


@pytest.fixture(scope="module")
def rsr():
    import reference_scripts_report

    return reference_scripts_report


@pytest.fixture(scope="module")
def double_bin():
    sys.path.insert(0, str(ROOT / "tests" / "abi_double"))
    import build_double

    return build_double.build()


def test_convergence_sweep_driver(rsr, double_bin, tmp_path):
    """scripts/convergence_sweep.py with its own options restricting the grid: every run it plans returns 0, the
    merged table has one row per run, and the rows are the reference's (analysis/data/convergence-results.csv)
    to the printed digits for the implicit schemes."""
    work = rsr.checkout(REFERENCE, tmp_path / "checkout", double_bin)
    res, _ = rsr.run_script(work, "convergence_sweep.py",
                            ["--nprocs", "4", "--nel", "10", "20", "--r", "1", "2", "--dt", "0.05", "0.01", "--schemes",
                             "theta-0.5", "theta-1.0", "newmark-0.25", "newmark-0.00"], env=rsr.TIGHT_CG, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    build = work / "build"
    runlog = list(csv.DictReader((build / "convergence-runlog.csv").open()))
    # 3 implicit schemes x 2 Nel x 2 R x 2 dt + the CFL-safe explicit runs (scripts/convergence_sweep.py:139-160)
    assert len(runlog) == 24 + 4 and all(r["returncode"] == "0" for r in runlog)
    rows, unmatched, _ = rsr.compare_convergence([build / "convergence-results.csv"])
    assert len(rows) == len(runlog) and unmatched == 0
    implicit = [r for r in rows if r["class"].startswith("implicit")]
    assert len(implicit) == 24 and all(r["dev"] <= 1e-6 for r in implicit), [r for r in implicit if r["dev"] > 1e-6]
    stable = [r for r in rows if r["class"].startswith("explicit (")]
    assert stable and all(r["dev"] <= 1e-4 for r in stable)
    # the driver's per-run logs hold the launcher's note and the program's output
    assert "Simulation completed" in (build / "convergence-logs" / "theta-0.5_Nel20_R2_dt0.01.out").read_text()
    assert "starting 1" in (build / "convergence-logs" / "theta-0.5_Nel20_R2_dt0.01.err").read_text()


def test_dissipation_dispersion_sweep_driver(rsr, double_bin, tmp_path):
    """scripts/dissipation_dispersion_sweep.py (Nel = 60, R = 1, T = 5 as in the reference's table, three time
    steps): the driver finds run-R1-N60x60-dt…/energy.csv, error.csv and probe.csv under the folder names it
    predicts (scripts/dissipation_dispersion_sweep.py:334-351) and its summary rows are the table's."""
    work = rsr.checkout(REFERENCE, tmp_path / "checkout", double_bin)
    res, _ = rsr.run_script(work, "dissipation_dispersion_sweep.py",
                            ["--nprocs", "4", "--dt", "0.15", "0.1", "0.05", "--schemes", "theta-0.5", "theta-1.0",
                             "newmark-0.25"], env=rsr.TIGHT_CG, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    build = work / "build"
    rows, unmatched, _ = rsr.compare_dissdisp(build / "dissdisp-results.csv")
    assert len(rows) == 9 and unmatched == 0
    for r in rows:
        assert r["dev"] is not None and r["dev"] <= 5e-6, r
    # theta = 1/2 and Newmark 1/4 conserve the energy: the ratio of the 6-digit energy.csv values is exactly 1
    assert all(r["energy_ratio"][0] == 1.0 for r in rows if r["key"][0] != "theta-1.0")
    for sub in ("dissdisp-energy-series", "dissdisp-error-series", "dissdisp-probe-series"):
        assert len(list((build / sub).glob("*.csv"))) == 9
    series = list(csv.DictReader((build / "dissdisp-probe-series" / "newmark-0.25_dt0.05.csv").open()))
    t, steps = 0.0, 0
    while t < 5.0:  # src/WaveNewmark.cpp:407-410: the accumulated float time decides the step count (101 here)
        t += 0.05
        steps += 1
    assert len(series) == steps and math.isfinite(float(series[-1]["u_probe"]))
