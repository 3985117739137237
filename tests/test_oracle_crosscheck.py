"""Second, independent restatement of the reference's time stepping (numpy + scipy, direct sparse
solves, pure-Python expression evaluation) against the C oracle, on the paths no reference artefact
pins: time-dependent forcing, inhomogeneous time-dependent Dirichlet data, variable wave speed, the
explicit Newmark boundary recurrence, P1 and P2.  Two separate implementations of the same reading
(src/WaveNewmark.cpp:56-278,290-390; src/WaveTheta.cpp:56-339) have to agree to solver
precision; the C oracle runs with a tight CG (reduce 1e-13).

Nothing here shares code with oracle/wave_oracle.c or the library: mesh, first-touch DoF numbering,
shape functions, quadrature tables, boundary treatment and the schemes are written again below."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import oracle as O
from wavegpu.problems import problem

TIGHT = dict(reduce=1e-13, tol=1e-30)


# ---- muParser subset -> Python (scalar evaluation) -----------------------------------------------
def compile_expr(block):
    text = block["Function expression"].replace("^", "**").replace("&&", " and ").replace("||", " or ")
    text = text.replace("if(", "_if(")
    env = {"pi": math.pi, "_if": lambda c, a, b: a if c else b}
    for name in ("sin", "cos", "exp", "sqrt", "tanh", "cosh", "sinh", "tan", "log"):
        env[name] = getattr(math, name)
    for item in filter(None, (s.strip() for s in block.get("Function constants", "").split(","))):
        k, v = item.split("=")
        env[k.strip()] = float(eval(v, {"pi": math.pi, "PI": math.pi}))
    code = compile(text, "<expr>", "eval")
    timed = "t" in block.get("Variable names", "")  # the reference's substring test

    def f(x, y, t=0.0):
        return float(eval(code, env, {"x": x, "y": y, "t": t if timed else 0.0}))

    return f


# ---- mesh, DoFs, element tables -------------------------------------------------------------------
def build_space(nx, ny, box, r):
    x0, x1, y0, y1 = box
    vid = lambda i, j: j * (nx + 1) + i
    vxy = [(x0 + i * (x1 - x0) / nx, y0 + j * (y1 - y0) / ny) for j in range(ny + 1) for i in range(nx + 1)]
    cells = []
    for j in range(ny):
        for i in range(nx):
            q0, q1, q2, q3 = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
            cells += [(q0, q1, q2), (q3, q2, q1)]
    vdof, ldof, pts, cell_dofs = {}, {}, [], []
    for c in cells:  # first touch: vertices in local order, then lines v0v1, v1v2, v2v0
        d = []
        for v in c:
            if v not in vdof:
                vdof[v] = len(pts)
                pts.append(vxy[v])
            d.append(vdof[v])
        if r == 2:
            for a, b in ((c[0], c[1]), (c[1], c[2]), (c[2], c[0])):
                key = (min(a, b), max(a, b))
                if key not in ldof:
                    ldof[key] = len(pts)
                    pts.append((0.5 * (vxy[a][0] + vxy[b][0]), 0.5 * (vxy[a][1] + vxy[b][1])))
                d.append(ldof[key])
        cell_dofs.append(d)
    pts = np.array(pts)
    eps = 1e-12 * max(x1 - x0, y1 - y0)
    on_b = (np.abs(pts[:, 0] - x0) < eps) | (np.abs(pts[:, 0] - x1) < eps) | (np.abs(pts[:, 1] - y0) < eps) | \
           (np.abs(pts[:, 1] - y1) < eps)
    return [vxy[v] for v in range(len(vxy))], cells, cell_dofs, pts, np.flatnonzero(on_b)


def shape(r, xi, eta):
    l0, l1, l2 = 1.0 - xi - eta, xi, eta
    if r == 1:
        return np.array([l0, l1, l2]), np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    phi = np.array([l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l0 * l1, 4 * l1 * l2, 4 * l2 * l0])
    g0, g1, g2 = np.array([-1.0, -1.0]), np.array([1.0, 0.0]), np.array([0.0, 1.0])
    grad = np.array([(4 * l0 - 1) * g0, (4 * l1 - 1) * g1, (4 * l2 - 1) * g2, 4 * (l0 * g1 + l1 * g0),
                     4 * (l1 * g2 + l2 * g1), 4 * (l2 * g0 + l0 * g2)])
    return phi, grad


def quadrature(n):
    if n == 2:  # 2x2 Gauss x Gauss-Jacobi(1,0), collapsed (deal.II >= 9.4 QGaussSimplex<2>(2))
        s6, s3 = math.sqrt(6.0), math.sqrt(3.0)
        out = []
        for sign in (-1.0, 1.0):
            for e, w in (((4 - s6) / 10, (9 + s6) / 72), ((4 + s6) / 10, (9 - s6) / 72)):
                out.append(((1 - e) * (1 + sign / s3) / 2, e, w))
        return out
    s = math.sqrt(15.0)
    a, b = (6 - s) / 21, (6 + s) / 21
    wa, wb = (155 - s) / 2400, (155 + s) / 2400
    return [(1 / 3, 1 / 3, 9 / 80)] + [(p, q, wa) for p, q in ((1 - 2 * a, a), (a, 1 - 2 * a), (a, a))] + \
           [(p, q, wb) for p, q in ((1 - 2 * b, b), (b, 1 - 2 * b), (b, b))]


class Space:
    def __init__(self, params):
        nel = [int(v) for v in params["Nel"].split(",")]
        self.nx, self.ny = (nel[0], nel[0]) if len(nel) == 1 else nel
        nums = [float(v) for v in params["Geometry"].replace("[", " ").replace("]", " ").replace("x", " ")
                .replace(",", " ").split()]
        self.r = int(params["R"])
        self.vxy, self.cells, self.cell_dofs, self.pts, self.bdofs = build_space(self.nx, self.ny, nums, self.r)
        self.n = len(self.pts)
        self.q = quadrature(self.r + 1)
        self.fn = {k: compile_expr(params[k]) for k in ("C", "F", "U0", "V0", "G", "DGDT")}
        self.M, self.K = self._matrices()

    def _geometry(self, c):
        v0, v1, v2 = (np.array(self.vxy[v]) for v in c)
        J = np.column_stack([v1 - v0, v2 - v0])
        return v0, J, abs(np.linalg.det(J)), np.linalg.inv(J)

    def _matrices(self):
        rows, cols, mv, kv = [], [], [], []
        for c, d in zip(self.cells, self.cell_dofs):
            v0, J, det, Jinv = self._geometry(c)
            Me = np.zeros((len(d), len(d)))
            Ke = np.zeros_like(Me)
            for xi, eta, w in self.q:
                phi, gref = shape(self.r, xi, eta)
                grad = gref @ Jinv  # rows: physical gradient of each shape function
                xq = v0 + J @ np.array([xi, eta])
                c2 = self.fn["C"](xq[0], xq[1], 0.0) ** 2
                Me += np.outer(phi, phi) * w * det
                Ke += c2 * (grad @ grad.T) * w * det
            for a, da in enumerate(d):
                for b, db in enumerate(d):
                    rows.append(da); cols.append(db); mv.append(Me[a, b]); kv.append(Ke[a, b])
        shape_ = (self.n, self.n)
        return sp.csr_matrix((mv, (rows, cols)), shape=shape_), sp.csr_matrix((kv, (rows, cols)), shape=shape_)

    def load(self, weights):  # sum_k w_k F(t_k) as a load vector
        out = np.zeros(self.n)
        for c, d in zip(self.cells, self.cell_dofs):
            v0, J, det, _ = self._geometry(c)
            for xi, eta, w in self.q:
                phi, _ = shape(self.r, xi, eta)
                xq = v0 + J @ np.array([xi, eta])
                fv = sum(wk * self.fn["F"](xq[0], xq[1], tk) for tk, wk in weights)
                out[d] += fv * phi * w * det
        return out

    def at_dofs(self, name, t=0.0, idx=None):
        idx = range(self.n) if idx is None else idx
        return np.array([self.fn[name](self.pts[i, 0], self.pts[i, 1], t) for i in idx])

    def value_at_centre(self, u):
        """u_h at the centre of the box: first cell that contains the point."""
        xs, ys = [q[0] for q in self.vxy], [q[1] for q in self.vxy]
        px, py = 0.5 * (min(xs) + max(xs)), 0.5 * (min(ys) + max(ys))
        for c, d in zip(self.cells, self.cell_dofs):
            v0, J, _, Jinv = self._geometry(c)
            xi, eta = Jinv @ (np.array([px, py]) - v0)
            if xi >= -1e-12 and eta >= -1e-12 and xi + eta <= 1 + 1e-12:
                phi, _ = shape(self.r, xi, eta)
                return float(phi @ u[d])
        raise AssertionError("centre not found")

    def solve_with_bc(self, A, rhs, values):
        """apply_boundary_values + exact solve: boundary rows become d0 e_i."""
        A = A.tolil(copy=True)
        rhs = rhs.copy()
        d0 = abs(A[0, 0])
        for i, v in zip(self.bdofs, values):
            A.rows[i], A.data[i] = [i], [d0]
            rhs[i] = v * d0
        return spla.spsolve(A.tocsc(), rhs)


def run_newmark(params, steps):
    S = Space(params)
    dt, beta, gamma = float(params["Dt"]), float(params["Beta"]), float(params["Gamma"])
    u, v = S.at_dofs("U0"), S.at_dofs("V0")
    g = lambda t: S.at_dofs("G", t, S.bdofs)
    a = S.solve_with_bc(S.M, S.load([(0.0, 1.0)]) - S.K @ u, (g(dt) - 2 * g(0.0) + g(-dt)) / dt ** 2)
    A = S.M + beta * dt * dt * S.K
    t = 0.0
    for _ in range(steps):
        t += dt
        z = u + dt * v + dt * dt * (0.5 - beta) * a
        rhs = -(S.K @ z) + S.load([(t, 1.0)])
        if beta > 1e-12:
            bc = (g(t) - z[S.bdofs]) / (beta * dt * dt)
        else:
            bc = (g(t) - 2 * g(t - dt) + g(t - 2 * dt)) / dt ** 2
        an = S.solve_with_bc(A, rhs, bc)
        u = z + dt * dt * beta * an
        v = v + dt * ((1 - gamma) * a + gamma * an)
        a = an
    return u, v, a


def run_theta(params, steps):
    S = Space(params)
    dt, th = float(params["Dt"]), float(params["Theta"])
    u, v = S.at_dofs("U0"), S.at_dofs("V0")
    Au = S.M + (th * dt) ** 2 * S.K
    t = 0.0
    for _ in range(steps):
        t += dt
        F = S.load([(t, th), (t - dt, 1 - th)])
        rhs_u = S.M @ u - dt * dt * th * (1 - th) * (S.K @ u) + dt * (S.M @ v) + th * dt * dt * F
        un = S.solve_with_bc(Au, rhs_u, S.at_dofs("G", t, S.bdofs))
        rhs_v = S.M @ v - dt * (1 - th) * (S.K @ u) - dt * th * (S.K @ un) + dt * F
        v = S.solve_with_bc(S.M, rhs_v, S.at_dofs("DGDT", t, S.bdofs))
        u = un
    return u, v, None


def oracle_state(params, scheme, steps):
    o = O.Oracle.from_params(params)
    o.set_cg(**TIGHT)
    if scheme == "newmark":
        o.newmark_init(float(params["Dt"]), float(params["Beta"]), float(params["Gamma"]))
    else:
        o.theta_init(float(params["Dt"]), float(params["Theta"]))
    for _ in range(steps):
        (o.newmark_step if scheme == "newmark" else o.theta_step)()
    return o, o.vector(O.Oracle.U), o.vector(O.Oracle.V), o.vector(O.Oracle.A) if scheme == "newmark" else None


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


VARIABLE_C = {"Function constants": "a=0.3", "Function expression": "1.0 + a*sin(2*pi*x)*cos(pi*y)",
              "Variable names": "x, y, t"}
CASES = [
    # time-dependent forcing (Ricker source widened so that a coarse mesh resolves it)
    ("ricker-wavelet", "newmark", dict(Nel="10, 8", R=1, Dt="0.01",
                                      F={"Function constants": "xs=0.5, ys=0.5, f0=20.0, sigma=0.1",
                                         "Function expression": "((1 - 2*(pi*f0*(t - 1/f0))^2) * exp(-(pi*f0*(t - 1/f0))^2))"
                                                                " * exp(-((x-xs)^2 + (y-ys)^2) / (2*sigma^2))",
                                         "Variable names": "x, y, t"})),
    ("ricker-wavelet", "theta", dict(Nel="6", R=2, Dt="0.01", Theta="1.0",
                                    F={"Function constants": "xs=0.5, ys=0.5, f0=20.0, sigma=0.1",
                                       "Function expression": "((1 - 2*(pi*f0*(t - 1/f0))^2) * exp(-(pi*f0*(t - 1/f0))^2))"
                                                              " * exp(-((x-xs)^2 + (y-ys)^2) / (2*sigma^2))",
                                       "Variable names": "x, y, t"})),
    ("dumping-wave", "theta", dict(Nel="8", R=1, Dt="0.02", Theta="0.5")),
    ("square-pulsing", "newmark", dict(Nel="7, 9", R=2, Dt="0.02")),
    # inhomogeneous, time-dependent Dirichlet data
    ("sine-membrane", "theta", dict(Nel="12, 6", R=1)),
    ("sine-membrane", "newmark", dict(Nel="9, 6", R=2)),
    ("oscillating-boundary", "newmark", dict(Nel="10", R=1, Dt="0.001", Beta="0.0")),  # explicit: a-BC recurrence
    ("oscillating-boundary", "theta", dict(Nel="6", R=2, Theta="1.0")),
    # variable wave speed
    ("gaussian-pulse", "newmark", dict(Nel="10", R=1, Dt="0.01", C=VARIABLE_C)),
    ("gaussian-pulse", "theta", dict(Nel="5, 7", R=2, Dt="0.01", C=VARIABLE_C)),
]


@pytest.mark.parametrize("name,scheme,over", CASES, ids=[f"{c[0]}-{c[1]}-R{c[2]['R']}" for c in CASES])
def test_independent_restatement_agrees_with_the_oracle(name, scheme, over):
    p = problem(name, **over)
    steps = 8
    u, v, a = (run_newmark if scheme == "newmark" else run_theta)(p, steps)
    o, ou, ov, oa = oracle_state(p, scheme, steps)
    assert len(u) == o.n
    assert np.abs(ou).max() > 0 or np.abs(ov).max() > 0  # the case exercises something
    assert rel(u, ou) < 2e-9, rel(u, ou)
    assert rel(v, ov) < 2e-9, rel(v, ov)
    if a is not None:
        assert rel(a, oa) < 2e-9, rel(a, oa)
    # compute_and_log_energy and log_point_probe on the independent state (src/WaveEquationBase.cpp:148-206)
    S = Space(p)
    E = 0.5 * (v @ (S.M @ v) + u @ (S.K @ u))
    assert abs(E - o.energy()) <= 1e-8 * max(abs(E), 1e-300)
    assert abs(S.value_at_centre(u) - o.probe()) <= 2e-9 * max(np.abs(ou).max(), 1e-300)


def test_independent_numbering_pattern_and_matrices():
    """The cell walk above gives the oracle's DoF numbering, sparsity pattern and element integrals."""
    p = problem("gaussian-pulse", Nel="5, 4", R=2, C=VARIABLE_C)
    S = Space(p)
    o = O.Oracle.from_params(p)
    sx, sy = o.support_points()
    assert np.allclose(S.pts[:, 0], sx, atol=1e-15) and np.allclose(S.pts[:, 1], sy, atol=1e-15)
    assert np.array_equal(S.bdofs, o.boundary_dofs())
    rowptr, col = o.csr()
    S.M.sort_indices()
    assert np.array_equal(S.M.indptr, rowptr) and np.array_equal(S.M.indices, col)
    S.K.sort_indices()
    assert rel(S.M.data, o.values(O.Oracle.M)) < 1e-13
    assert rel(S.K.data, o.values(O.Oracle.K)) < 1e-12
