"""Further GPU checks: size-independent properties at BASELINE sizes for P2 and the theta scheme, API
semantics (wave_run == repeated wave_step, re-init, error codes), solver options."""
import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import WaveSolver, api, problem

pytestmark = pytest.mark.gpu


def rel(a, b):
    den = np.abs(b).max()
    return np.abs(np.asarray(a) - np.asarray(b)).max() / (den if den > 0 else 1.0)


def test_full_size_properties_p2():
    """Nel=2048, R=2 (16.8 M DoFs, 193 M nnz): K symmetric, constants in its kernel, mass = area,
    BC rows of the system matrix are d0 * e_i, Newmark(1/4,1/2) conserves the discrete energy."""
    p = problem("standing-mode-wsol", Nel="2048", R=2, Dt="0.002")
    g = WaveSolver(p, "newmark")
    n = g.n
    assert n == (2 * 2048 + 1) ** 2 and g.nnz_local == 46 * 2048 ** 2 + 16 * 2048 + 1
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Kx, Ky = g.spmv(api.MAT_K, x), g.spmv(api.MAT_K, y)
    assert abs(y @ Kx - x @ Ky) < 1e-9 * abs(y @ Kx)
    ones = np.ones(n)
    assert np.abs(g.spmv(api.MAT_K, ones)).max() < 1e-8
    assert abs(ones @ g.spmv(api.MAT_M, ones) - 1.0) < 1e-12
    bd = g.boundary_dofs()
    assert bd.size == 8 * 2048
    e = np.zeros(n)
    e[bd] = 1.0
    Se = g.spmv(api.MAT_SYS1, e)
    d0 = Se[bd[0]]
    assert d0 > 0 and np.allclose(Se[bd], d0, rtol=0, atol=0)
    g.init()
    g.run(2)
    e0 = g.energy()
    g.run(10)
    assert abs(g.energy() - e0) < 1e-5 * e0
    g.close()


def test_theta_full_size_energy():
    p = problem("standing-mode-wsol", Nel="1024", R=1, Dt="0.01", Theta="0.5")
    g = WaveSolver(p, "theta")
    g.init()
    g.run(2)
    e0 = g.energy()
    done, its, nrm, tot = g.run(10)
    assert done == 10 and its[0] > 0 and its[1] > 0
    assert abs(g.energy() - e0) < 1e-5 * e0
    g.close()


@pytest.mark.parametrize("scheme", ["newmark", "theta"])
def test_run_equals_repeated_step(scheme):
    p = problem("sine-membrane", Nel="30, 10", R=2)
    a, b = WaveSolver(p, scheme), WaveSolver(p, scheme)
    a.init()
    b.init()
    for _ in range(9):
        a.step()
    done, its, nrm, tot = b.run(9)
    assert done == 9 and a.time == b.time
    assert np.array_equal(a.vector(api.VEC_U), b.vector(api.VEC_U))  # deterministic assembly and reductions
    a.close()
    b.close()


def test_reinit_restarts_the_run():
    p = problem("gaussian-pulse", Nel="24", R=2)
    g = WaveSolver(p, "newmark")
    g.init()
    for _ in range(4):
        g.step()
    u1 = g.vector(api.VEC_U)
    g.init()
    for _ in range(4):
        g.step()
    assert np.array_equal(g.vector(api.VEC_U), u1)
    g.close()


def test_divergence_and_nonconvergence_codes():
    p = problem("gaussian-pulse", Nel="32", Dt="0.05", Beta="0.0")  # explicit, far above the CFL bound
    g = WaveSolver(p, "newmark")
    g.init()
    with pytest.raises(api.WaveError) as ei:
        g.run(2000)
    assert ei.value.code == -6 and g.step_no < 2000
    g.close()
    p = problem("standing-mode-wsol", Nel="64", Dt="0.05")
    g = WaveSolver(p, "newmark", cg=dict(maxit=3))
    with pytest.raises(api.WaveError) as ei:
        g.init()
        for _ in range(3):
            g.step()
    assert ei.value.code == -5
    g.close()


def test_identity_preconditioner_matches_oracle():
    p = problem("standing-mode-wsol", Nel="18", R=2, Dt="0.01")
    cg = dict(precond=1)
    o = O.Oracle.from_params(p)
    o.set_cg(precond=1)
    o.newmark_init(0.01, 0.25, 0.5)
    g = WaveSolver(p, "newmark", cg=cg)
    g.init()
    for _ in range(6):
        o.newmark_step()
        its, _ = g.step()
        assert its == o.iterations()
    assert rel(g.vector(api.VEC_U), o.vector(O.Oracle.U)) < 1e-10
    g.close()


def test_forcing_every_step_flag_is_equivalent():
    """F = 0.0 with the forcing kernels forced to run every step: the load vector is exactly zero, and
    assembly is a deterministic row gather, so the two contexts agree bit for bit."""
    p = problem("standing-mode-wsol", Nel="20", Dt="0.02", Theta="0.5")
    a = WaveSolver(p, "theta")
    b = WaveSolver(p, "theta", flags=api.FLAG_FORCING_EVERY_STEP)
    a.init()
    b.init()
    a.run(8)
    b.run(8)
    assert np.array_equal(a.vector(api.VEC_V), b.vector(api.VEC_V))
    assert np.array_equal(a.vector(api.VEC_U), b.vector(api.VEC_U))
    assert b.launch_count() > a.launch_count()
    a.close()
    b.close()


def test_set_vector_roundtrip_and_norms():
    p = problem("standing-mode-wsol", Nel="13, 7", R=2)
    g = WaveSolver(p, "newmark")
    rng = np.random.default_rng(2)
    u, v = rng.standard_normal(g.n), rng.standard_normal(g.n)
    g.set_vector(api.VEC_U, u)
    g.set_vector(api.VEC_V, v)
    assert np.array_equal(g.vector(api.VEC_U), u) and np.array_equal(g.vector(api.VEC_V), v)
    nu, nv = g.norms()
    assert nu == pytest.approx(np.linalg.norm(u), rel=1e-13) and nv == pytest.approx(np.linalg.norm(v), rel=1e-13)
    g.close()


def test_reference_default_case_full_run():
    """BASELINE configs[0]: parameters/sine-membrane.json as shipped (theta = 0.5, Nel 180 x 60 on
    [0,3]x[0,1], Dt = 0.05, T = 60: 1201 steps, inhomogeneous time-dependent Dirichlet data) -- the whole
    run on the GPU against the whole run of the oracle: energy series, iteration counts, final vectors."""
    p = problem("sine-membrane")
    out = O.run(p, "theta", log_every=50)
    assert out["steps"] == 1201
    g = WaveSolver(p, "theta")
    g.init()
    k, same_its, total = 0, 0, 0
    t = 0.0
    for step in range(1, out["steps"] + 1):
        its, _ = g.step()
        if step % 50 == 0:
            s, tt, E = out["energy"][k]
            assert s == step
            assert abs(g.energy() - E) <= 1e-8 * abs(E)
            oi = out["iterations"][k][2:]
            same_its += int(tuple(its) == tuple(oi))
            total += 1
            k += 1
    assert same_its >= total - 1  # round-off may move a stopping iteration by one, very rarely
    o = out["oracle"]
    assert rel(g.vector(api.VEC_U), o.vector(O.Oracle.U)) < 1e-8
    assert rel(g.vector(api.VEC_V), o.vector(O.Oracle.V)) < 1e-8
    g.close()
