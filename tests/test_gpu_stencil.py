"""The table-driven ("stencil", matrix-free) rows of the SpMV against the SELL rows and the oracle.

With a constant wave speed on the structured mesh almost every row of M, K and M + sK is a translate of
one representative row per DoF kind; wave_setup detects those rows numerically and k_spmv serves them
from a table in shared memory (csrc/kernels.cuh).  The sums run over the same entries in the same order;
the table values differ from the assembled ones only by the round-off of dx = x_{i+1} - x_i."""
import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import WaveSolver, api, problem

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


@pytest.mark.parametrize("nel,r", [("64", 1), ("96, 40", 1), ("300", 1), ("150", 2), ("200, 90", 2), ("512", 2)])
def test_stencil_spmv_matches_sell_and_oracle(nel, r):
    p = problem("standing-mode-wsol", Nel=nel, R=r, Dt="0.01")
    o = O.Oracle.from_params(p)
    a = WaveSolver(p, "theta")
    b = WaveSolver(p, "theta", flags=api.FLAG_NO_STENCIL)
    ia, ib = a.operator_info(), b.operator_info()
    assert ib["stencil_rows"] == 0 and ib["sell_nnz"] == b.nnz_local
    # (on small P2 meshes the windows of 1024 rows mix the DoF kinds and fewer than half of the rows sit in
    # single-kind slices: the stencil path then stays off -- covered by test_small_meshes...)
    assert ia["stencil_rows"] >= 0.5 * a.n and ia["stencil_rows"] + ia["sell_rows"] == a.n
    assert ia["spmv_bytes"] < ib["spmv_bytes"]
    rng = np.random.default_rng(11)
    x = rng.standard_normal(o.n)
    for gid, oid in ((api.MAT_M, O.Oracle.M), (api.MAT_K, O.Oracle.K)):
        ya, yb, yo = a.spmv(gid, x), b.spmv(gid, x), o.spmv(oid, x)
        assert rel(ya, yb) < 1e-12
        assert rel(ya, yo) < 1e-13
    for gid in (api.MAT_SYS1, api.MAT_SYS2):
        assert rel(a.spmv(gid, x), b.spmv(gid, x)) < 1e-12
    a.close()
    b.close()


def test_stencil_share_grows_with_the_mesh():
    for nel, r, share in (("512", 1, 0.95), ("256", 2, 0.75), ("1024", 2, 0.93)):
        g = WaveSolver(problem("standing-mode-wsol", Nel=nel, R=r), "newmark")
        info = g.operator_info()
        assert info["stencil_rows"] >= share * g.n, (nel, r, info, g.n)
        g.close()


def test_variable_wave_speed_keeps_sell_rows():
    p = problem("standing-mode-wsol", Nel="64", R=2,
                C={"Function constants": "", "Function expression": "1.0 + 0.25*sin(2*pi*x)*sin(2*pi*y)",
                   "Variable names": "x, y, t"})
    g = WaveSolver(p, "newmark")
    assert g.operator_info()["stencil_rows"] == 0
    g.close()


def test_small_meshes_have_no_stencil_rows():
    g = WaveSolver(problem("standing-mode-wsol", Nel="6", R=2), "newmark")
    assert g.operator_info()["stencil_rows"] == 0
    g.close()


@pytest.mark.parametrize("name,scheme,over", [
    ("standing-mode-wsol", "newmark", dict(Nel="40", R=1, Dt="0.01")),
    ("standing-mode-wsol", "newmark", dict(Nel="160", R=2, Dt="0.002")),
    ("standing-mode-wsol", "theta", dict(Nel="200, 120", R=2, Dt="0.002", Theta="0.5")),
    ("ricker-wavelet", "theta", dict(Nel="40", Theta="1.0")),
    ("sine-membrane", "newmark", dict(Nel="60, 20")),
])
@pytest.mark.parametrize("stencil", [True, False])
def test_time_stepping_parity_both_operators(name, scheme, over, stencil):
    """Oracle parity (solution vectors within 1e-10, identical CG iteration counts) with the table-driven
    rows and with every row kept in SELL form."""
    p = problem(name, **over)
    o = O.Oracle.from_params(p)
    g = WaveSolver(p, scheme, flags=0 if stencil else api.FLAG_NO_STENCIL)
    assert (g.operator_info()["stencil_rows"] > 0) == stencil
    dt = float(p["Dt"])
    if scheme == "newmark":
        o.newmark_init(dt, float(p["Beta"]), float(p["Gamma"]))
    else:
        o.theta_init(dt, float(p["Theta"]))
    g.init()
    for _ in range(12):
        (o.newmark_step if scheme == "newmark" else o.theta_step)()
        its, _ = g.step()
        assert its == o.iterations()
    assert rel(g.vector(api.VEC_U), o.vector(O.Oracle.U)) < 1e-10
    assert rel(g.vector(api.VEC_V), o.vector(O.Oracle.V)) < 1e-10
    assert abs(g.energy() - o.energy()) <= 1e-10 * max(abs(o.energy()), 1e-300)
    g.close()


def test_stencil_full_size_properties():
    """Nel=2048, R=2 with the table-driven rows: K symmetric, constants in its kernel, mass = area, and the
    same step as the SELL operator to round-off."""
    p = problem("standing-mode-wsol", Nel="2048", R=2, Dt="0.002")
    g = WaveSolver(p, "newmark")
    n = g.n
    assert g.operator_info()["stencil_rows"] > 0.95 * n
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Kx, Ky = g.spmv(api.MAT_K, x), g.spmv(api.MAT_K, y)
    assert abs(y @ Kx - x @ Ky) < 1e-9 * abs(y @ Kx)
    ones = np.ones(n)
    assert np.abs(g.spmv(api.MAT_K, ones)).max() < 1e-8
    assert abs(ones @ g.spmv(api.MAT_M, ones) - 1.0) < 1e-12
    g.init()
    g.run(3)
    u = g.vector(api.VEC_U)
    g.close()
    b = WaveSolver(p, "newmark", flags=api.FLAG_NO_STENCIL)
    b.init()
    b.run(3)
    assert rel(u, b.vector(api.VEC_U)) < 1e-10
    b.close()
