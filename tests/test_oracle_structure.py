"""Structural facts the reference implies, checked on the oracle: DoF and
non-zero counts, row lengths of the sparsity pattern, boundary sizes, exactness of the element integrals
for constant c, quadrature degrees."""
import ctypes as C
from math import factorial as f

import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import problem


@pytest.mark.parametrize("N", [3, 8, 17])
def test_counts_p1(N):
    o = O.Oracle.from_params(problem("standing-mode-wsol", Nel=N, R=1))
    assert o.n == (N + 1) ** 2 and o.nnz == 7 * N * N + 6 * N + 1 and o.nb == 4 * N and o.ncells == 2 * N * N
    rp, _ = o.csr()
    lens = np.diff(rp)
    assert set(lens.tolist()) <= {7, 5, 4, 3}
    assert (lens == 7).sum() == (N - 1) ** 2  # interior vertices


@pytest.mark.parametrize("N", [2, 5, 12])
def test_counts_p2(N):
    o = O.Oracle.from_params(problem("standing-mode-wsol", Nel=N, R=2))
    assert o.n == (2 * N + 1) ** 2 and o.nnz == 46 * N * N + 16 * N + 1 and o.nb == 8 * N
    rp, _ = o.csr()
    lens = np.diff(rp)
    assert set(lens.tolist()) <= {19, 9, 12, 6}
    assert (lens == 19).sum() == (N - 1) ** 2               # interior vertices
    assert (lens == 9).sum() == 3 * N * N - 2 * N + 2       # interior edges + the two 2-cell corner vertices


def test_rectangular_mesh_counts():
    o = O.Oracle.from_params(problem("sine-membrane", Nel="9, 4", R=2))
    assert o.n == (2 * 9 + 1) * (2 * 4 + 1) and o.ncells == 2 * 9 * 4 and o.nb == 4 * (9 + 4)


def test_p1_element_matrices_are_the_analytic_ones():
    """QGaussSimplex(r+1) integrates M and K exactly for constant c:
    M_e = A/12 (1 + delta_ij), K_e = A G^T G on the right triangles of a uniform mesh."""
    N = 4
    o = O.Oracle.from_params(problem("standing-mode-wsol", Nel=N, R=1))
    rp, col = o.csr()
    M, K = o.values(O.Oracle.M), o.values(O.Oracle.K)
    h = 1.0 / N
    area = h * h / 2
    cd = o.cell_dofs()
    # an interior vertex belongs to 6 cells: diagonal of M = 6 * 2 * A/12 = A
    interior = [i for i in range(o.n) if rp[i + 1] - rp[i] == 7]
    for i in interior:
        row = slice(rp[i], rp[i + 1])
        d = M[row][col[row] == i][0]
        assert d == pytest.approx(area, rel=1e-14)
        assert M[row].sum() == pytest.approx(h * h, rel=1e-14)          # row sum = patch area / 3 * 3
        assert K[row].sum() == pytest.approx(0.0, abs=1e-13)           # constants in the kernel
        assert K[row][col[row] == i][0] == pytest.approx(4.0, rel=1e-14)  # 5-point-like stencil: 4 on the diagonal
    assert M.sum() == pytest.approx(1.0, rel=1e-14)                       # 1^T M 1 = |domain|
    assert cd.min() == 0 and cd.max() == o.n - 1


def test_p2_mass_is_exact():
    o = O.Oracle.from_params(problem("sine-membrane", Nel="6, 3", R=2))
    assert o.values(O.Oracle.M).sum() == pytest.approx(3.0, rel=1e-13)    # box [0,3] x [0,1]
    x, y = o.support_points()
    M = o.values(O.Oracle.M)
    # the P2 space contains quadratics: x^T M y = integral of (x^2)(y) over the box = 9 * 1/2
    v1, v2 = x * x, y
    assert v1 @ o.spmv(O.Oracle.M, v2) == pytest.approx(9.0 * 0.5, rel=1e-12)
    # and K reproduces the Dirichlet form: integral of grad(x^2/2 + y) . grad(x y) = int (x*y + x) = 9/4 + 9/2
    assert (0.5 * x * x + y) @ o.spmv(O.Oracle.K, x * y) == pytest.approx(9.0 / 4 + 9.0 / 2, rel=1e-12)
    assert M.min() < 0  # P2 mass matrices have negative vertex-vertex couplings


@pytest.mark.parametrize("n1d,nq,deg", [(2, 4, 3), (3, 7, 5), (4, 15, 7)])
def test_quadrature_degree(n1d, nq, deg):
    xi, eta, w = np.zeros(16), np.zeros(16), np.zeros(16)
    dp = C.POINTER(C.c_double)
    got = O.lib().oracle_get_quadrature(n1d, xi.ctypes.data_as(dp), eta.ctypes.data_as(dp), w.ctypes.data_as(dp))
    assert got == nq and w[:nq].sum() == pytest.approx(0.5, rel=1e-14)
    for a in range(deg + 1):
        for b in range(deg + 1 - a):
            exact = f(a) * f(b) / f(a + b + 2)
            assert (w[:nq] * xi[:nq] ** a * eta[:nq] ** b).sum() == pytest.approx(exact, abs=1e-15)
    # and not exact one degree higher
    a = deg + 1
    assert abs((w[:nq] * xi[:nq] ** a).sum() - f(a) / f(a + 2)) > 1e-12
