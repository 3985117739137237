"""Pin the CPU oracle (oracle/wave_oracle.c) to the reference's own shipped result tables.

Known answers: analysis/data/convergence-results.csv and dissdisp-results.csv of the reference,
extracted by tests/golden/extract_golden.py (recipes: scripts/convergence_sweep.py:165-179,
scripts/dissipation_dispersion_sweep.py:179-198, 249-330).

The reference solves with AMG-CG stopped at 1e-6 residual reduction; its preconditioner is not
reproduced (north star: Jacobi).  The discretisation itself is pinned by solving tightly
(reduce 1e-13), where all rows (P1 and P2) agree to the 7 printed digits; with the reference's own
stopping rule (1e-6) the agreement is what two different preconditioners allow (~1e-5)."""
import json
from pathlib import Path

import pytest

from oracle import oracle as O
from wavegpu.problems import problem

GOLD = Path(__file__).resolve().parent / "golden"
CONV = json.loads((GOLD / "convergence_rows.json").read_text())
DISS = json.loads((GOLD / "dissdisp_rows.json").read_text())
TIGHT = dict(reduce=1e-13, tol=1e-30)


def _params(row, **extra):
    kw = dict(Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"])
    for k in ("Theta", "Beta", "Gamma"):
        if row.get(k) is not None:
            kw[k] = row[k]
    kw.update(extra)
    return problem("standing-mode-wsol", **kw)


def _rel(a, b):
    return abs(a - b) / abs(b)


@pytest.mark.parametrize("row", CONV, ids=lambda r: f"L{r['line']}-{r['scheme']}-N{r['Nel']}-r{r['R']}-dt{r['Dt']}")
def test_convergence_row(row):
    out = O.run(_params(row), row["scheme"], cg=TIGHT)
    _, _, rl2, rh1 = out["final_errors"]
    # both degrees within the 7 printed digits: the error rules are deal.II's own (7-point for r=1,
    # 15-point Witherden-Vincent for r=2), the per-cell float rounding is replicated
    tol_l2 = tol_h1 = 1e-6
    # marginally stable explicit runs amplify solver-level differences
    if row["rel_H1"] > 3 * row["rel_L2"] and row["rel_H1"] > 0.5:
        tol_l2, tol_h1 = 1e-4, 1e-4
    assert _rel(rl2, row["rel_L2"]) < tol_l2
    assert _rel(rh1, row["rel_H1"]) < tol_h1


def test_convergence_reference_stopping_rule():
    """Same rows with ReductionControl(10000, 1e-12, 1e-6) (src/WaveNewmark.cpp:256)."""
    for row in [r for r in CONV if r["Nel"] == 20 and r["Dt"] in ("0.05", "0.01")]:
        out = O.run(_params(row), row["scheme"])
        _, _, rl2, rh1 = out["final_errors"]
        loose = 2e-2 if (row["rel_H1"] > 3 * row["rel_L2"] and row["rel_H1"] > 0.5) else 2e-4
        assert _rel(rl2, row["rel_L2"]) < loose and _rel(rh1, row["rel_H1"]) < loose


@pytest.mark.parametrize("row", DISS, ids=lambda r: f"L{r['line']}-{r['scheme']}-dt{r['Dt']}")
def test_dissdisp_row(row):
    kind, val = row["scheme"].split("-")
    scheme = "theta" if kind == "theta" else "newmark"
    extra = {"Theta": val} if kind == "theta" else {"Beta": val, "Gamma": "0.5"}
    p = problem("standing-mode-wsol", Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"], **extra)
    out = O.run(p, scheme, log_every=1, cg=TIGHT)
    # energy.csv carries 6 significant digits (src/WaveEquationBase.cpp:166)
    E = [float("%.6g" % e[2]) for e in out["energy"]]
    ratio = E[-1] / E[0]
    if row["energy_ratio"] == 1.0:
        assert abs(ratio - 1.0) < 2e-6
    else:
        assert _rel(ratio, row["energy_ratio"]) < 5e-6
    rl2 = [e[4] for e in out["error"]]
    assert _rel(max(rl2), row["max_rel_L2"]) < 5e-6
    assert _rel(rl2[-1], row["final_rel_L2"]) < 5e-6
    assert _rel(out["error"][-1][5], row["final_rel_H1"]) < 5e-6


def test_step_counts_of_shipped_files():
    """while (time < T) float accumulation (src/WaveNewmark.cpp:407-410)."""
    expect = {"dumping-wave": 858, "five-modes-wsol": 4801, "gaussian-pulse": 481, "oscillating-boundary": 601,
              "ricker-wavelet": 572, "sine-membrane-likedeal2": 320, "sine-membrane": 1201, "square-bump": 6001,
              "square-pulsing": 572, "standing-mode-wsol": 6001, "traveling-square-bump": 320,
              "two-modes-wsol": 572}
    for name, steps in expect.items():
        p = problem(name)
        t, dt, T, n = 0.0, float(p["Dt"]), float(p["T"]), 0
        while t < T:
            t += dt
            n += 1
        assert n == steps, name


def test_theta_half_equals_newmark_quarter():
    """theta=1/2 == Newmark(1/4,1/2) for f=g=0 (SURVEY 8c analytic invariant)."""
    a = O.run(problem("standing-mode-wsol", Nel=12, Dt=0.02, T=0.5, Theta=0.5), "theta", cg=TIGHT)
    b = O.run(problem("standing-mode-wsol", Nel=12, Dt=0.02, T=0.5), "newmark", cg=TIGHT)
    import numpy as np

    ua, ub = a["oracle"].vector(0), b["oracle"].vector(0)
    assert np.abs(ua - ub).max() < 1e-9 * np.abs(ub).max()


def test_multigrid_preconditioner_in_the_oracle():
    """precond=2 (geometric V-cycle): an order of magnitude fewer CG iterations than
    Jacobi, the same discrete solution (both stop at 1e-6 residual reduction), and the reference's
    known answers with a tight solve."""
    import numpy as np

    p = problem("standing-mode-wsol", Nel=64, R=2, Dt=0.05, T=1.0)
    oj = O.Oracle.from_params(p)
    om = O.Oracle.from_params(p)
    om.set_cg(precond=2)
    oj.newmark_init(0.05, 0.25, 0.5)
    om.newmark_init(0.05, 0.25, 0.5)
    ij = im = 0
    for _ in range(4):
        oj.newmark_step()
        om.newmark_step()
        ij += oj.iterations()[0]
        im += om.iterations()[0]
    assert im * 8 < ij
    assert np.abs(om.vector(0) - oj.vector(0)).max() < 1e-6 * np.abs(oj.vector(0)).max()
    row = next(r for r in CONV if r["scheme"] == "newmark" and r["Nel"] == 20 and r["R"] == 1
               and r["Dt"] == "0.05" and r["Beta"] == 0.25)
    out = O.run(_params(row), "newmark", cg=dict(precond=2, **TIGHT))
    assert _rel(out["final_errors"][2], row["rel_L2"]) < 2e-6
    assert _rel(out["final_errors"][3], row["rel_H1"]) < 2e-6
