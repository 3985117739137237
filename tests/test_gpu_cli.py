"""GPU tests of the drop-in surface: the main-newmark / main-theta executables run a parameter file
and must write the reference's artefacts (folder naming, energy.csv / error.csv / iterations.csv /
probe.csv / convergence.csv formats) with the oracle's numbers."""
import csv
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import problem
from wavegpu.problems import write_json

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "nmpde-wave-equation_b200" / "bin"


def _fmt6(x):  # default ostream formatting: 6 significant digits (src/WaveEquationBase.cpp:166)
    return float("%.6g" % x)


@pytest.mark.parametrize("exe,scheme,folder", [
    ("main-newmark", "newmark", "run-R1-N12x12-dt0_05-T1-gamma0_5-beta0_25"),
    ("main-theta", "theta", "run-R1-N12x12-dt0_05-T1-theta1"),
])
def test_cli_run_writes_reference_artefacts(exe, scheme, folder, tmp_path):
    (tmp_path / "build").mkdir()
    (tmp_path / "parameters").mkdir()
    p = problem("standing-mode-wsol", Nel="12", R="1", Dt="0.05", T="1.0", Theta="1.0", Save_Solution=False,
                Log_Every=2, Print_Every=5)
    write_json(tmp_path / "parameters" / "conv-params.json", p)
    r = subprocess.run([str(BIN / exe), "../parameters/conv-params.json"], cwd=tmp_path / "build",
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    run_dir = tmp_path / "results" / f"{scheme}-conv-params" / folder
    assert run_dir.is_dir(), list((tmp_path / "results").rglob("*"))
    out = O.run(p, scheme, log_every=2)
    assert f"Simulation completed: {out['steps']} steps" in r.stdout

    rows = list(csv.reader((run_dir / "energy.csv").open()))
    assert rows[0] == ["timestep", "time", "energy"]
    assert len(rows) - 1 == len(out["energy"])
    for row, (step, t, E) in zip(rows[1:], out["energy"]):
        assert int(row[0]) == step
        assert float(row[1]) == _fmt6(t)
        assert float(row[2]) == pytest.approx(_fmt6(E), rel=2e-6)

    rows = list(csv.reader((run_dir / "error.csv").open()))
    assert rows[0] == ["timestep", "time", "L2_error", "H1_error", "rel_L2_error", "rel_H1_error"]
    for k, (row, e) in enumerate(zip(rows[1:], out["error"])):
        assert int(row[0]) == e[0]
        # sticky std::scientific: from the second row on the time column is scientific too (App. A.8)
        assert ("e" in row[1]) == (k > 0)
        assert np.allclose([float(v) for v in row[2:]], e[2:], rtol=2e-6)

    rows = list(csv.reader((run_dir / "iterations.csv").open()))
    assert rows[0] == ["timestep", "time", "iterations_1", "iterations_2"]
    assert [int(r_[2]) for r_ in rows[1:]] == [it[2] for it in out["iterations"]]
    assert [int(r_[3]) for r_ in rows[1:]] == [it[3] for it in out["iterations"]]

    rows = list(csv.reader((run_dir / "probe.csv").open()))
    assert rows[0] == ["timestep", "time", "u_probe"]
    assert np.allclose([float(r_[2]) for r_ in rows[1:]], [pr[2] for pr in out["probe"]], rtol=1e-9, atol=1e-14)

    conv = list(csv.reader((tmp_path / "results" / f"{scheme}-conv-params" / "convergence.csv").open()))
    assert conv[0][:4] == ["h", "N_el_x", "N_el_y", "r"] and len(conv) == 2
    assert float(conv[1][10]) == pytest.approx(out["final_errors"][2], rel=2e-6)
    if scheme == "newmark":
        assert conv[1][7:10] == ["N/A", "0.250000", "0.500000"]
        assert (run_dir / "parameters.json").exists()  # NMPDE_PARAM_FILE is set by main-newmark only
    else:
        assert conv[1][7:10] == ["1.000000", "N/A", "N/A"]


def test_cli_divergence_is_reported_and_exit_code_zero(tmp_path):
    (tmp_path / "build").mkdir()
    (tmp_path / "parameters").mkdir()
    # explicit Newmark far above the CFL bound: check_divergence must stop the loop (exit code 0)
    p = problem("gaussian-pulse", Nel="32", Dt="0.05", T="30.0", Beta="0.0", Save_Solution=False, Enable_Logging=False)
    write_json(tmp_path / "parameters" / "blowup.json", p)
    r = subprocess.run([str(BIN / "main-newmark"), "../parameters/blowup.json"], cwd=tmp_path / "build",
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0
    assert "Divergence detected at step" in r.stdout
