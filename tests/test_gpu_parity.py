"""GPU parity tests: libwavegpu.so (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): sparsity pattern and DoF numbering bit-exact; matrices, solution
vectors and the energy / error series within 1e-10 relative (observed ~1e-13); CG converged to the
same tolerance with the same iteration counts."""
import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import WaveSolver, problem
from wavegpu import api

pytestmark = pytest.mark.gpu

TIGHT = dict(reduce=1e-13, tol=1e-30)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


MESHES = [("1", 1), ("1", 2), ("2, 1", 2), ("1, 3", 1), ("3", 1), ("4", 2), ("7, 5", 1), ("6, 9", 2),
          ("33", 1), ("20, 31", 2)]


@pytest.mark.parametrize("nel,r", MESHES)
def test_pattern_numbering_matrices(nel, r):
    p = problem("standing-mode-wsol", Nel=nel, R=r)
    o = O.Oracle.from_params(p)
    g = WaveSolver(p, "theta")
    assert g.n == o.n and g.nnz_local == o.nnz
    rp, col, M = g.csr(api.MAT_M)
    _, _, K = g.csr(api.MAT_K)
    orp, ocol = o.csr()
    assert np.array_equal(rp, orp), "row pointers differ"
    assert np.array_equal(col, ocol), "column indices differ"
    assert rel(M, o.values(O.Oracle.M)) < 1e-14
    assert rel(K, o.values(O.Oracle.K)) < 1e-13
    sx, sy = g.support_points()
    osx, osy = o.support_points()
    assert np.array_equal(sx, osx) and np.array_equal(sy, osy)
    assert np.array_equal(g.boundary_dofs(), o.boundary_dofs())
    g.close()


@pytest.mark.parametrize("r", [1, 2])
def test_assembly_is_bitwise_reproducible(r):
    """Row-gather assembly (no atomics): M, K and a forcing load vector of two separate contexts are equal
    bit for bit."""
    p = problem("square-pulsing", Nel="23, 17", R=r,
                C={"Function constants": "", "Function expression": "1.0 + 0.3*sin(3*x)*cos(2*y)",
                   "Variable names": "x, y, t"})
    a, b = WaveSolver(p, "newmark"), WaveSolver(p, "newmark")
    for mid in (api.MAT_M, api.MAT_K, api.MAT_SYS1):
        assert np.array_equal(a.csr(mid)[2], b.csr(mid)[2])
    a.init()
    b.init()
    a.step()
    b.step()
    assert np.array_equal(a.vector(api.VEC_A), b.vector(api.VEC_A))
    a.close()
    b.close()


def test_variable_wave_speed_assembly():
    p = problem("standing-mode-wsol", Nel="9, 7", R=2,
                C={"Function constants": "a=0.3", "Function expression": "1.0 + a*sin(2*pi*x)*cos(pi*y)",
                   "Variable names": "x, y, t"})
    o = O.Oracle.from_params(p)
    g = WaveSolver(p, "newmark")
    _, _, K = g.csr(api.MAT_K)
    assert rel(K, o.values(O.Oracle.K)) < 1e-13
    g.close()


@pytest.mark.parametrize("nel,r", [("17", 1), ("12, 19", 2), ("300", 1), ("150", 2)])
def test_spmv(nel, r):
    p = problem("standing-mode-wsol", Nel=nel, R=r)
    o = O.Oracle.from_params(p)
    g = WaveSolver(p, "theta")
    rng = np.random.default_rng(7)
    x = rng.standard_normal(o.n)
    for gid, oid in ((api.MAT_M, O.Oracle.M), (api.MAT_K, O.Oracle.K)):
        y = g.spmv(gid, x)
        yo = o.spmv(oid, x)
        assert rel(y, yo) < 1e-13
    g.close()


@pytest.mark.parametrize("nel,r,scheme", [("24", 1, "newmark"), ("16", 2, "theta")])
def test_cg_same_iterations_and_solution(nel, r, scheme):
    p = problem("standing-mode-wsol", Nel=nel, R=r, Dt="0.05")
    o = O.Oracle.from_params(p)
    g = WaveSolver(p, scheme)
    if scheme == "newmark":
        o.newmark_init(0.05, 0.25, 0.5)
        g.init()
        o.newmark_step()
    else:
        o.theta_init(0.05, 0.5)
        g.init()
        o.theta_step()
    # the oracle's last BC-modified system matrix is matrix S; solve S x = b from x0 with both
    _, _, S1 = g.csr(api.MAT_SYS1 if scheme == "newmark" else api.MAT_SYS2)
    So = o.values(O.Oracle.S)
    assert rel(S1, So) < 1e-13
    rng = np.random.default_rng(3)
    b = rng.standard_normal(o.n)
    bd = o.boundary_dofs()
    b[bd] = 0.0
    x0 = np.zeros(o.n)
    xo, ito = o.cg(O.Oracle.S, x0, b)
    xg, itg = g.cg(api.MAT_SYS1 if scheme == "newmark" else api.MAT_SYS2, x0, b)
    assert itg == ito
    assert rel(xg, xo) < 1e-10
    g.close()


def _march(p, scheme, nsteps, cg=None, check_every=1):
    o = O.Oracle.from_params(p)
    g = WaveSolver(p, scheme, cg=cg)
    if cg:
        o.set_cg(**cg)
    dt = float(p["Dt"])
    if scheme == "newmark":
        o.newmark_init(dt, float(p["Beta"]), float(p["Gamma"]))
    else:
        o.theta_init(dt, float(p["Theta"]))
    g.init()
    worst = 0.0
    its_equal = True
    for s in range(nsteps):
        (o.newmark_step if scheme == "newmark" else o.theta_step)()
        its, nrm = g.step()
        oi = o.iterations()
        its_equal &= (its == oi)
        if (s + 1) % check_every == 0 or s == nsteps - 1:
            worst = max(worst, rel(g.vector(api.VEC_U), o.vector(O.Oracle.U)),
                        rel(g.vector(api.VEC_V), o.vector(O.Oracle.V)))
            assert abs(nrm[0] - o.norm(O.Oracle.U)) <= 1e-10 * max(o.norm(O.Oracle.U), 1e-300)
    return o, g, worst, its_equal


@pytest.mark.parametrize("name,scheme,over", [
    ("standing-mode-wsol", "newmark", dict(Nel="20", R=1, Dt="0.02")),
    ("standing-mode-wsol", "newmark", dict(Nel="12", R=2, Dt="0.01")),
    ("standing-mode-wsol", "newmark", dict(Nel="16", R=1, Dt="0.005", Beta="0.0")),
    ("standing-mode-wsol", "theta", dict(Nel="20", R=1, Dt="0.02", Theta="0.5")),
    ("standing-mode-wsol", "theta", dict(Nel="10", R=2, Dt="0.02", Theta="1.0")),
    ("standing-mode-wsol", "theta", dict(Nel="16", R=1, Dt="0.002", Theta="0.0")),
    ("sine-membrane", "theta", dict(Nel="30, 10")),
    ("sine-membrane", "newmark", dict(Nel="30, 10")),
    ("sine-membrane", "newmark", dict(Nel="24, 8", Dt="0.005", Beta="0.0")),
    ("oscillating-boundary", "newmark", dict(Nel="16", R=2)),
    ("ricker-wavelet", "theta", dict(Nel="20", Theta="1.0")),
    ("ricker-wavelet", "newmark", dict(Nel="14", R=2)),
    ("dumping-wave", "theta", dict(Nel="16")),
    ("square-pulsing", "newmark", dict(Nel="16")),
    ("gaussian-pulse", "newmark", dict(Nel="24")),
    ("traveling-square-bump", "newmark", dict(Nel="30, 10")),
    ("square-bump", "theta", dict(Nel="20")),
])
@pytest.mark.parametrize("cg_path", ["auto", "three-kernel"])
def test_time_stepping_parity(name, scheme, over, cg_path, monkeypatch):
    """Reference stopping rule ReductionControl(10000, 1e-12, 1e-6): same iteration counts and
    solution vectors within 1e-10 relative of the oracle's -- with the default CG path (the cooperative
    kernel K6f at these sizes) and with the three-kernel iteration."""
    if cg_path == "three-kernel":
        monkeypatch.setenv("WAVE_CG_FUSED", "0")
    p = problem(name, **over)
    o, g, worst, its_equal = _march(p, scheme, 25)
    assert worst < 1e-10, worst
    assert its_equal
    assert abs(g.energy() - o.energy()) <= 1e-10 * max(abs(o.energy()), 1e-300)
    g.close()


def test_energy_and_error_series():
    p = problem("standing-mode-wsol", Nel="20", R=1, Dt="0.05", T="1.0", Theta="1.0")
    out = O.run(p, "theta", log_every=1)
    g = WaveSolver(p, "theta")
    g.init()
    for (step, t, E), err in zip(out["energy"], out["error"]):
        g.step()
        assert abs(g.energy() - E) <= 1e-10 * abs(E)
        ge = g.errors()
        # float per-cell rounding (src/WaveEquationBase.cpp:384) is replicated, so L2 is tight too;
        # the H1 part differentiates the exact solution by centred differences with h = 1e-8
        # ([deal.II] AutoDerivativeFunction), which amplifies 1-ulp libm differences by 1/(2h)
        assert np.allclose([ge[0], ge[2]], [err[2], err[4]], rtol=1e-9, atol=0)
        assert np.allclose([ge[1], ge[3]], [err[3], err[5]], rtol=2e-7, atol=0)
    g.close()


def test_error_norms_p2():
    p = problem("two-modes-wsol", Nel="12", R=2, Dt="0.01")
    o, g, worst, _ = _march(p, "newmark", 5)
    ge, oe = g.errors(), o.errors()
    assert np.allclose([ge[0], ge[2]], [oe[0], oe[2]], rtol=1e-9, atol=0)
    assert np.allclose([ge[1], ge[3]], [oe[1], oe[3]], rtol=2e-7, atol=0)
    assert abs(g.probe(0.5, 0.5) - o.probe()) < 1e-12
    g.close()


def test_golden_rows_through_the_gpu():
    """Known answers of the reference (analysis/data/convergence-results.csv) via libwavegpu."""
    import json
    from pathlib import Path

    rows = json.loads((Path(__file__).parent / "golden" / "convergence_rows.json").read_text())
    picked = [r for r in rows if r["Nel"] == 20 and r["Dt"] in ("0.05", "0.01") and r["R"] == 1]
    assert picked
    for row in picked:
        kw = dict(Nel=row["Nel"], R=row["R"], Dt=row["Dt"], T=row["T"])
        for k in ("Theta", "Beta", "Gamma"):
            if row.get(k) is not None:
                kw[k] = row[k]
        p = problem("standing-mode-wsol", **kw)
        g = WaveSolver(p, row["scheme"], cg=TIGHT)
        g.init()
        t, dt, T = 0.0, float(p["Dt"]), float(p["T"])
        while t < T:
            t += dt
            g.step()
        e = g.errors()
        assert abs(e[2] - row["rel_L2"]) / row["rel_L2"] < 2e-6
        assert abs(e[3] - row["rel_H1"]) / row["rel_H1"] < 2e-6
        g.close()


def test_full_size_properties():
    """BASELINE config 2 at full size (Nel=1024, R=1): properties that need no oracle run --
    symmetry of K via x.Ky = y.Kx, row sums of K = 0 (constants are in its kernel), mass = area,
    energy conservation of Newmark(1/4,1/2)."""
    p = problem("standing-mode-wsol", Nel="1024", R=1, Dt="0.01")
    g = WaveSolver(p, "newmark")
    n = g.n
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Kx, Ky = g.spmv(api.MAT_K, x), g.spmv(api.MAT_K, y)
    assert abs(y @ Kx - x @ Ky) < 1e-9 * abs(y @ Kx)
    ones = np.ones(n)
    assert np.abs(g.spmv(api.MAT_K, ones)).max() < 1e-9
    assert abs(ones @ g.spmv(api.MAT_M, ones) - 1.0) < 1e-12
    g.init()
    g.run(3)
    e0 = g.energy()
    g.run(20)
    assert abs(g.energy() - e0) < 1e-5 * e0  # CG stops at 1e-6 reduction
    g.close()


def test_step_host_matches_resident_path():
    p = problem("standing-mode-wsol", Nel="20", R=2, Dt="0.02")
    a = WaveSolver(p, "newmark")
    b = WaveSolver(p, "newmark")
    a.init()
    b.init()
    u, v, acc = b.vector(api.VEC_U), b.vector(api.VEC_V), b.vector(api.VEC_A)
    for _ in range(5):
        a.step()
        b.step_host(u, v, acc)
    assert np.array_equal(a.vector(api.VEC_U), u)  # row-gather assembly: two contexts agree bit for bit
    a.close()
    b.close()


def test_errors_are_loud():
    with pytest.raises(api.WaveError):
        WaveSolver(problem("standing-mode-wsol", R=3, Nel=4), "newmark")
    bad = problem("standing-mode-wsol", Nel=4)
    bad["U0"]["Function expression"] = "sin(pi*x"
    with pytest.raises(api.WaveError) as ei:
        WaveSolver(bad, "newmark")
    assert ei.value.code == -2
