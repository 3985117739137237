"""CPU tests of the host side of the drop-in -- the product's ParameterReader / WaveEquationBase / WaveNewmark /
WaveTheta / cli / launcher sources, unchanged -- linked with the TEST DOUBLE of the C ABI
(tests/abi_double/wave_abi_on_oracle.cpp: the entry points the host classes call, on the CPU oracle) instead of
libwavegpu.so.  The executables under test live in tests/_build/ and are test infrastructure: the product's
executables (nmpde-wave-equation_b200/bin/) have no CPU path and are covered on the GPU by tests/test_gpu_cli.py
with the same assertions.  What is checked here is everything above the ABI: the reference's artefacts
(folder naming, energy.csv / error.csv / iterations.csv / probe.csv / convergence.csv formats),
the time loop's step count, exit codes, the launcher's command line."""
import csv
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import problem
from wavegpu.problems import write_json

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "tools" / "mpirun-shim"


@pytest.fixture(scope="module")
def double_bin():
    sys.path.insert(0, str(ROOT / "tests" / "abi_double"))
    import build_double

    return build_double.build()


def _fmt6(x):  # default ostream formatting: 6 significant digits (src/WaveEquationBase.cpp:166)
    return float("%.6g" % x)


def _layout(tmp_path, name, params):
    (tmp_path / "build").mkdir()
    (tmp_path / "parameters").mkdir()
    write_json(tmp_path / "parameters" / name, params)
    return tmp_path / "build"


@pytest.mark.parametrize("exe,scheme,folder,over", [
    ("main-newmark", "newmark", "run-R1-N12x12-dt0_05-T1-gamma0_5-beta0_25", dict(Nel="12", R="1")),
    ("main-theta", "theta", "run-R1-N12x12-dt0_05-T1-theta1", dict(Nel="12", R="1")),
    ("main-newmark", "newmark", "run-R2-N6x9-dt0_05-T1-gamma0_5-beta0_25", dict(Nel="6, 9", R="2")),
])
def test_host_classes_write_the_reference_artefacts(double_bin, exe, scheme, folder, over, tmp_path):
    p = problem("standing-mode-wsol", Dt="0.05", T="1.0", Theta="1.0", Save_Solution=False, Log_Every=2,
                Print_Every=5, **over)
    build = _layout(tmp_path, "conv-params.json", p)
    r = subprocess.run([str(double_bin / exe), "../parameters/conv-params.json"], cwd=build, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    run_dir = tmp_path / "results" / f"{scheme}-conv-params" / folder
    assert run_dir.is_dir(), list((tmp_path / "results").rglob("*"))
    out = O.run(p, scheme, log_every=2)
    assert f"Simulation completed: {out['steps']} steps" in r.stdout

    rows = list(csv.reader((run_dir / "energy.csv").open()))
    assert rows[0] == ["timestep", "time", "energy"]
    assert len(rows) - 1 == len(out["energy"])
    for row, (step, t, E) in zip(rows[1:], out["energy"]):
        assert int(row[0]) == step and float(row[1]) == _fmt6(t) and float(row[2]) == _fmt6(E)

    rows = list(csv.reader((run_dir / "error.csv").open()))
    assert rows[0] == ["timestep", "time", "L2_error", "H1_error", "rel_L2_error", "rel_H1_error"]
    for k, (row, e) in enumerate(zip(rows[1:], out["error"])):
        assert int(row[0]) == e[0]
        assert ("e" in row[1]) == (k > 0)  # sticky std::scientific from the second row on (App. A.8)
        assert np.allclose([float(v) for v in row[2:]], e[2:], rtol=1e-6)

    rows = list(csv.reader((run_dir / "iterations.csv").open()))
    assert rows[0] == ["timestep", "time", "iterations_1", "iterations_2"]
    assert [(int(r_[2]), int(r_[3])) for r_ in rows[1:]] == [(it[2], it[3]) for it in out["iterations"]]

    rows = list(csv.reader((run_dir / "probe.csv").open()))
    assert rows[0] == ["timestep", "time", "u_probe"]
    assert np.allclose([float(r_[2]) for r_ in rows[1:]], [pr[2] for pr in out["probe"]], rtol=1e-9, atol=1e-14)

    conv = list(csv.reader((tmp_path / "results" / f"{scheme}-conv-params" / "convergence.csv").open()))
    assert conv[0] == ["h", "N_el_x", "N_el_y", "r", "dt", "T", "method", "theta", "beta", "gamma",
                       "rel_L2_error_final", "rel_H1_error_final", "elapsed_time_s"] and len(conv) == 2
    assert float(conv[1][10]) == pytest.approx(out["final_errors"][2], rel=1e-6)
    assert float(conv[1][11]) == pytest.approx(out["final_errors"][3], rel=1e-6)
    if scheme == "newmark":
        assert conv[1][6:10] == ["newmark-conv-params", "N/A", "0.250000", "0.500000"]
        assert (run_dir / "parameters.json").exists()  # NMPDE_PARAM_FILE is set by main-newmark only
    else:
        assert conv[1][6:10] == ["theta-conv-params", "1.000000", "N/A", "N/A"]
        assert not (run_dir / "parameters.json").exists()

    # a second run appends to convergence.csv without repeating the header (scripts/convergence_sweep.py:308-321)
    r = subprocess.run([str(double_bin / exe), "../parameters/conv-params.json"], cwd=build, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0
    conv = list(csv.reader((tmp_path / "results" / f"{scheme}-conv-params" / "convergence.csv").open()))
    assert len(conv) == 3 and conv[1][:12] == conv[2][:12]


def test_divergence_stops_the_loop_with_exit_code_zero(double_bin, tmp_path):
    p = problem("gaussian-pulse", Nel="16", Dt="0.05", T="30.0", Beta="0.0", Save_Solution=False, Enable_Logging=False)
    build = _layout(tmp_path, "blowup.json", p)
    r = subprocess.run([str(double_bin / "main-newmark"), "../parameters/blowup.json"], cwd=build, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0
    assert "Divergence detected at step" in r.stdout
    out = O.run(p, "newmark")
    assert f"Divergence detected at step {out['diverged']}," in r.stdout


def test_log_every_zero_writes_no_series(double_bin, tmp_path):
    """The convergence driver's parameter files (scripts/convergence_sweep.py:165-179): Enable Logging false,
    Log Every 0 -- no energy/error/probe/iterations files, only convergence.csv."""
    p = problem("standing-mode-wsol", Nel="8", Dt="0.1", T="0.5", Save_Solution=False, Enable_Logging=False, Log_Every=0)
    build = _layout(tmp_path, "conv-params.json", p)
    r = subprocess.run([str(double_bin / "main-theta"), "../parameters/conv-params.json"], cwd=build,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    files = sorted(f.name for f in (tmp_path / "results").rglob("*") if f.is_file())
    assert files == ["convergence.csv"]


def test_save_solution_writes_vtu_of_the_host_vectors(double_bin, tmp_path):
    from wavegpu import cell_dofs
    from wavegpu.vtu import read_vtu

    p = problem("standing-mode-wsol", Nel="6, 4", R="2", Dt="0.05", T="0.15", Save_Solution=True, Log_Every=1)
    build = _layout(tmp_path, "vtu.json", p)
    r = subprocess.run([str(double_bin / "main-newmark"), "../parameters/vtu.json"], cwd=build, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = O.run(p, "newmark", log_every=1)
    steps = out["steps"]
    run_dir = next(d for d in (tmp_path / "results" / "newmark-vtu").iterdir() if d.is_dir())
    assert sorted(f.name for f in run_dir.glob("solution_*")) == sorted(
        [f"solution_{k:04d}.0.vtu" for k in range(steps + 1)] + [f"solution_{k:04d}.pvtu" for k in range(steps + 1)])
    pts, conn, offs, types, data = read_vtu(run_dir / f"solution_{steps:04d}.0.vtu")
    corner = cell_dofs(6, 4, 2)[:, :3].ravel()
    o = out["oracle"]
    assert np.array_equal(data["u"], o.vector(O.Oracle.U)[corner])
    assert np.array_equal(data["v"], o.vector(O.Oracle.V)[corner])


def test_launcher_command_line_of_the_sweep_drivers(double_bin, tmp_path):
    """`<launcher> -np 4 --bind-to core --map-by socket <binary> <file>` (scripts/convergence_sweep.py:182-210)
    through tools/mpirun-shim -> the product's wave-mpirun: placement options dropped, one process where no GPU is
    visible, the program's exit status returned (1 for an unreadable parameter file, src/main-newmark.cpp:92-97)."""
    p = problem("standing-mode-wsol", Nel="8", Dt="0.1", T="0.3", Save_Solution=False, Log_Every=0)
    build = _layout(tmp_path, "conv-params.json", p)
    cmd = [str(SHIM), "-np", "4", "--bind-to", "core", "--map-by", "socket", str(double_bin / "main-newmark")]
    r = subprocess.run(cmd + [str(tmp_path / "parameters" / "conv-params.json")], cwd=build, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Simulation completed: 3 steps") == 1
    assert (tmp_path / "results" / "newmark-conv-params" / "convergence.csv").exists()
    r = subprocess.run(cmd + [str(tmp_path / "parameters" / "missing.json")], cwd=build, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 1
