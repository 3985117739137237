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


@pytest.mark.parametrize("exe,scheme,name,over", [
    ("main-newmark", "newmark", "standing-mode-wsol", dict(Nel="24", R="2", Dt="0.02", T="0.5")),
    ("main-theta", "theta", "sine-membrane", dict(Nel="30, 10", T="2.0")),
])
def test_cli_two_ranks_write_the_one_rank_artefacts(exe, scheme, name, over, tmp_path):
    """`mpirun -np 2 main-...` of the reference -> `wave-mpirun -np 2 main-...`: two GPUs, strips of quad
    rows, rank 0 writes the files; every series equals the one-GPU run's (needs >= 2 GPUs)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = problem(name, Save_Solution=False, Log_Every=2, Print_Every=5, **over)
    series = {}
    for ranks in (1, 2):
        root = tmp_path / f"np{ranks}"
        (root / "build").mkdir(parents=True)
        (root / "parameters").mkdir()
        write_json(root / "parameters" / "case.json", p)
        r = subprocess.run([str(BIN / "wave-mpirun"), "-np", str(ranks), "--bind-to", "core", str(BIN / exe),
                            "../parameters/case.json"], cwd=root / "build", capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("Simulation completed") == 1  # rank 0 speaks for the run
        assert (f"{ranks} GPUs" in r.stdout) == (ranks > 1)
        run_dirs = [d for d in (root / "results" / f"{scheme}-case").iterdir() if d.is_dir()]
        assert len(run_dirs) == 1
        series[ranks] = {f.name: list(csv.reader(f.open())) for f in sorted(run_dirs[0].glob("*.csv"))}
    assert set(series[1]) == set(series[2]) and "energy.csv" in series[1] and "iterations.csv" in series[1]
    assert series[1]["iterations.csv"] == series[2]["iterations.csv"]
    # printed digits: energy 6, errors 7 (per-cell float sums, order differs between partitions), probe 11
    rtol = {"energy.csv": 2e-6, "error.csv": 2e-6, "probe.csv": 1e-9}
    for fname in series[1]:
        a, b = series[1][fname], series[2][fname]
        assert a[0] == b[0] and len(a) == len(b)
        for ra, rb in zip(a[1:], b[1:]):
            assert np.allclose([float(x) for x in ra], [float(x) for x in rb], rtol=rtol.get(fname, 0.0), atol=1e-14)


def test_cli_save_solution_writes_vtu(tmp_path):
    """"Save Solution": true (the reference's default): solution_NNNN.0.vtu + .pvtu at step 0 and after
    every step with u, v, u_exact and partitioning as point data (src/WaveEquationBase.cpp:330-365)."""
    from wavegpu import cell_dofs
    from wavegpu.vtu import read_vtu

    (tmp_path / "build").mkdir()
    (tmp_path / "parameters").mkdir()
    p = problem("standing-mode-wsol", Nel="6, 4", R="2", Dt="0.05", T="0.15", Save_Solution=True, Log_Every=1)
    write_json(tmp_path / "parameters" / "vtu.json", p)
    r = subprocess.run([str(BIN / "main-newmark"), "../parameters/vtu.json"], cwd=tmp_path / "build",
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = O.run(p, "newmark", log_every=1)
    steps = out["steps"]
    run_dir = next(d for d in (tmp_path / "results" / "newmark-vtu").iterdir() if d.is_dir())
    assert sorted(f.name for f in run_dir.glob("solution_*")) == sorted(
        [f"solution_{k:04d}.0.vtu" for k in range(steps + 1)] + [f"solution_{k:04d}.pvtu" for k in range(steps + 1)])
    pts, conn, offs, types, data = read_vtu(run_dir / f"solution_{steps:04d}.0.vtu")
    ncells = 2 * 6 * 4
    assert len(pts) == 3 * ncells and np.all(types == 5) and offs[-1] == 3 * ncells
    assert set(data) == {"u", "v", "u_exact", "partitioning"} and not data["partitioning"].any()
    corner = cell_dofs(6, 4, 2)[:, :3].ravel()
    o = out["oracle"]
    u, v = o.vector(O.Oracle.U), o.vector(O.Oracle.V)
    assert np.abs(data["u"] - u[corner]).max() <= 1e-10 * np.abs(u).max()
    assert np.abs(data["v"] - v[corner]).max() <= 1e-10 * np.abs(v).max()
    sx, sy = o.support_points()
    assert np.abs(pts[:, 0] - sx[corner]).max() < 1e-6 and np.abs(pts[:, 1] - sy[corner]).max() < 1e-6
    t_end = out["time"]
    exact = np.cos(np.sqrt(2.0) * np.pi * t_end) * np.sin(np.pi * sx[corner]) * np.sin(np.pi * sy[corner])
    assert np.abs(data["u_exact"] - exact).max() < 1e-12
    # step 0 holds the initial condition
    _, _, _, _, d0 = read_vtu(run_dir / "solution_0000.0.vtu")
    assert np.abs(d0["u"] - np.sin(np.pi * sx[corner]) * np.sin(np.pi * sy[corner])).max() < 1e-12


def test_convergence_sweep_recipe_through_the_launcher(tmp_path):
    """The recipe of the reference's scripts/convergence_sweep.py:165-210 against its own result table: a
    parameter file written the way `write_param_file` does (Nel, R, Dt, T, Save Solution / Enable Logging
    off, Log Every 0, scheme overrides), started as `<launcher> -np 1 <binary> <file>` from build/ with the
    MPI options the script may add, and `../results/<method>-<stem>/convergence.csv` read back -- rows of
    analysis/data/convergence-results.csv (tests/golden/convergence_rows.json) within their printed digits."""
    import json
    import os

    shim = ROOT / "tools" / "mpirun-shim"
    rows = json.loads((ROOT / "tests" / "golden" / "convergence_rows.json").read_text())
    picked = [r for r in rows if r["Nel"] == 20 and r["Dt"] == "0.01" and r["R"] in (1, 2)
              and r["scheme"] in ("theta", "newmark")]
    picked = [r for r in picked if (r.get("Theta") in (0.5, 1.0)) or (r.get("Beta") == 0.25)][:6]
    assert len(picked) >= 4
    build = tmp_path / "build"
    build.mkdir()
    (tmp_path / "parameters").mkdir()
    base = problem("standing-mode-wsol")
    env = dict(os.environ, WAVE_CG_REDUCE="1e-13", WAVE_CG_TOL="1e-30")
    for k, row in enumerate(picked):
        params = dict(base)
        params.update({"Nel": str(row["Nel"]), "R": str(row["R"]), "Dt": str(row["Dt"]), "T": str(row["T"]),
                       "Save Solution": False, "Enable Logging": False, "Log Every": 0})
        for key in ("Theta", "Beta", "Gamma"):
            if row.get(key) is not None:
                params[key] = str(row[key])
        pf = tmp_path / "parameters" / "convergence-params.json"
        write_json(pf, params)
        exe = BIN / ("main-theta" if row["scheme"] == "theta" else "main-newmark")
        cmd = [str(shim), "-np", "1", "--bind-to", "core", "--map-by", "socket", str(exe), str(pf)]
        r = subprocess.run(cmd, cwd=build, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        conv = tmp_path / "results" / f"{row['scheme']}-convergence-params" / "convergence.csv"
        lines = list(csv.reader(conv.open()))
        assert lines[0][:4] == ["h", "N_el_x", "N_el_y", "r"] and lines[0][-3:-1] == ["rel_L2_error_final",
                                                                                     "rel_H1_error_final"]
        last = lines[-1]
        assert int(last[1]) == row["Nel"] and int(last[3]) == row["R"]
        l2, h1 = float(last[-3]), float(last[-2])
        assert abs(l2 - row["rel_L2"]) <= 3e-6 * row["rel_L2"], (row, l2)
        assert abs(h1 - row["rel_H1"]) <= 3e-6 * row["rel_H1"], (row, h1)
