// mesh.h -- analytic structured-triangle mesh, closed-form DoF numbering, P1/P2 shape functions
// and quadrature rules, usable from host and device.
//
// Replaces (as data the kernels need, never stored as tables):
//   GridGenerator::subdivided_hyper_rectangle_with_simplices   src/WaveEquationBase.cpp:42-46
//   DoFHandler::distribute_dofs(FE_SimplexP(r)) on one rank    src/WaveEquationBase.cpp:90-91
//   FE_SimplexP(r) / QGaussSimplex(r+1)                        src/WaveEquationBase.cpp:78,82
//
// Mesh: vertices (i,j), 0<=i<=Nx, 0<=j<=Ny at (x0+i dx, y0+j dy).  Quads are walked j outer,
// i inner; quad (i,j) has q0=(i,j) q1=(i+1,j) q2=(i,j+1) q3=(i+1,j+1) and two cells
// T0={q0,q1,q2} (cell 2(j Nx+i)) and T1={q3,q2,q1} (cell 2(j Nx+i)+1).
//
// Numbering: deal.II's first touch in cell order -- per cell the unnumbered vertices (v0,v1,v2),
// then the unnumbered lines (v0v1, v1v2, v2v0).  On this mesh that gives the closed forms below.
// Every DoF introduced while walking quad row j forms one contiguous "block j":
//   P1: block 0 = vertex lines 0 and 1 interleaved (2(Nx+1) DoFs); block j>=1 = vertex line j+1.
//   P2: block 0 has 6Nx+3 DoFs (9 for quad 0, 6 per further quad); block j>=1 has 4Nx+2
//       (6 for quad 0, 4 per further quad).
// Edge entities of quad (i,j): B(i,j) bottom (i,j)-(i+1,j); L(i,j) left (i,j)-(i,j+1);
// D(i,j) diagonal (i+1,j)-(i,j+1).  Top of quad (i,j) is B(i,j+1), right is L(i+1,j).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define WV_HD __host__ __device__ __forceinline__
#else
#define WV_HD inline
#endif

namespace wv {

struct Mesh {
    int nx, ny, r;
    double x0, y0, dx, dy;
};

WV_HD int dofs_per_cell(int r) { return r == 1 ? 3 : 6; }
WV_HD int64_t n_dofs(const Mesh &m) {
    return m.r == 1 ? (int64_t)(m.nx + 1) * (m.ny + 1) : (int64_t)(2 * m.nx + 1) * (2 * m.ny + 1);
}
WV_HD int64_t n_cells(const Mesh &m) { return 2LL * m.nx * m.ny; }

// first DoF of block j (j may equal ny: one past the last block)
WV_HD int64_t block_start(const Mesh &m, int j) {
    if (j <= 0) return 0;
    if (m.r == 1) return (int64_t)(j + 1) * (m.nx + 1);
    return (int64_t)(6 * m.nx + 3) + (int64_t)(j - 1) * (4 * m.nx + 2);
}

// ---- forward maps entity -> global DoF ------------------------------------------------
WV_HD int64_t dof_V(const Mesh &m, int i, int j) {
    if (m.r == 1) {
        if (j == 0) return i < 2 ? i : 2 * i;
        if (j == 1) return i == 0 ? 2 : 2 * i + 1;
        return (int64_t)j * (m.nx + 1) + i;
    }
    if (j == 0) return i < 2 ? i : 9 + 6 * (int64_t)(i - 2);
    if (j == 1) return i == 0 ? 2 : (i == 1 ? 6 : 9 + 6 * (int64_t)(i - 2) + 3);
    const int64_t S = block_start(m, j - 1);
    return i == 0 ? S : (i == 1 ? S + 3 : S + 6 + 4 * (int64_t)(i - 2) + 1);
}
// the three edge maps exist for r == 2 only
WV_HD int64_t dof_B(const Mesh &m, int i, int j) {
    if (j == 0) return i == 0 ? 3 : 9 + 6 * (int64_t)(i - 1) + 1;
    if (j == 1) return i == 0 ? 7 : 9 + 6 * (int64_t)(i - 1) + 4;
    const int64_t S = block_start(m, j - 1);
    return i == 0 ? S + 4 : S + 6 + 4 * (int64_t)(i - 1) + 2;
}
WV_HD int64_t dof_L(const Mesh &m, int i, int j) {
    if (j == 0) return i == 0 ? 5 : (i == 1 ? 8 : 9 + 6 * (int64_t)(i - 2) + 5);
    const int64_t S = block_start(m, j);
    return i == 0 ? S + 2 : (i == 1 ? S + 5 : S + 6 + 4 * (int64_t)(i - 2) + 3);
}
WV_HD int64_t dof_D(const Mesh &m, int i, int j) {
    if (j == 0) return i == 0 ? 4 : 9 + 6 * (int64_t)(i - 1) + 2;
    const int64_t S = block_start(m, j);
    return i == 0 ? S + 1 : S + 6 + 4 * (int64_t)(i - 1);
}

// cell -> its dofs_per_cell global DoFs in deal.II's cell-local order
// (3 vertices, then line0=v0v1, line1=v1v2, line2=v2v0)
WV_HD void cell_dofs(const Mesh &m, int64_t cell, int64_t *d) {
    const int64_t q = cell >> 1;
    const int j = (int)(q / m.nx), i = (int)(q - (int64_t)j * m.nx);
    if ((cell & 1) == 0) {  // T0 = {q0,q1,q2}
        d[0] = dof_V(m, i, j); d[1] = dof_V(m, i + 1, j); d[2] = dof_V(m, i, j + 1);
        if (m.r == 2) { d[3] = dof_B(m, i, j); d[4] = dof_D(m, i, j); d[5] = dof_L(m, i, j); }
    } else {  // T1 = {q3,q2,q1}
        d[0] = dof_V(m, i + 1, j + 1); d[1] = dof_V(m, i, j + 1); d[2] = dof_V(m, i + 1, j);
        if (m.r == 2) { d[3] = dof_B(m, i, j + 1); d[4] = dof_D(m, i, j); d[5] = dof_L(m, i + 1, j); }
    }
}

// ---- internal (storage) numbering ---------------------------------------------------------------
// Vectors and matrix rows are stored in a permutation of the canonical numbering that keeps every
// block [block_start(j), block_start(j+1)) but lists its DoFs kind by kind (all vertices of the line,
// then the horizontal, vertical and diagonal edge DoFs, each in ascending i).  Neighbouring DoFs of
// one kind are then neighbours in memory, so the k-th column of 32 consecutive rows is a contiguous
// run: SpMV gathers coalesce for P2 as they do for P1.  For r == 1 the internal numbering is the
// canonical one.  The C ABI speaks canonical numbering only (wave_get_vector / wave_get_csr permute).
WV_HD int64_t idof_V(const Mesh &m, int i, int j) {
    if (m.r == 1) return dof_V(m, i, j);
    if (j == 0) return i;
    if (j == 1) return (int64_t)(m.nx + 1) + i;
    return block_start(m, j - 1) + i;
}
WV_HD int64_t idof_B(const Mesh &m, int i, int j) {
    if (j == 0) return 2 * (int64_t)(m.nx + 1) + i;
    if (j == 1) return 2 * (int64_t)(m.nx + 1) + m.nx + i;
    return block_start(m, j - 1) + (m.nx + 1) + i;
}
WV_HD int64_t idof_L(const Mesh &m, int i, int j) {
    if (j == 0) return 2 * (int64_t)(m.nx + 1) + 2 * (int64_t)m.nx + i;
    return block_start(m, j) + (m.nx + 1) + m.nx + i;
}
WV_HD int64_t idof_D(const Mesh &m, int i, int j) {
    if (j == 0) return 2 * (int64_t)(m.nx + 1) + 2 * (int64_t)m.nx + (m.nx + 1) + i;
    return block_start(m, j) + (m.nx + 1) + m.nx + (m.nx + 1) + i;
}
WV_HD void cell_dofs_internal(const Mesh &m, int64_t cell, int64_t *d) {
    const int64_t q = cell >> 1;
    const int j = (int)(q / m.nx), i = (int)(q - (int64_t)j * m.nx);
    if ((cell & 1) == 0) {
        d[0] = idof_V(m, i, j); d[1] = idof_V(m, i + 1, j); d[2] = idof_V(m, i, j + 1);
        if (m.r == 2) { d[3] = idof_B(m, i, j); d[4] = idof_D(m, i, j); d[5] = idof_L(m, i, j); }
    } else {
        d[0] = idof_V(m, i + 1, j + 1); d[1] = idof_V(m, i, j + 1); d[2] = idof_V(m, i + 1, j);
        if (m.r == 2) { d[3] = idof_B(m, i, j + 1); d[4] = idof_D(m, i, j); d[5] = idof_L(m, i + 1, j); }
    }
}

// affine geometry of a cell: origin (v0), and the diagonal Jacobian (sx, sy) with x = X0 + sx*xi,
// y = Y0 + sy*eta.  T0: (dx, dy); T1: (-dx, -dy).  |det J| = dx*dy for both.
WV_HD void cell_geometry(const Mesh &m, int64_t cell, double &X0, double &Y0, double &sx, double &sy) {
    const int64_t q = cell >> 1;
    const int j = (int)(q / m.nx), i = (int)(q - (int64_t)j * m.nx);
    // v1 - v0 and v2 - v0 are differences of the vertex positions x0 + i*dx (not dx itself),
    // the way a mesh stored as vertex coordinates yields them
    const double xa = m.x0 + i * m.dx, xb = m.x0 + (i + 1) * m.dx;
    const double ya = m.y0 + j * m.dy, yb = m.y0 + (j + 1) * m.dy;
    if ((cell & 1) == 0) { X0 = xa; Y0 = ya; sx = xb - xa; sy = yb - ya; }
    else { X0 = xb; Y0 = yb; sx = xa - xb; sy = ya - yb; }
}

// ---- entity enumeration -----------------------------------------------------------------
// Every DoF is one "entity slot" (j, i, kind) with kind 0=V, 1=B, 2=L, 3=D (kinds 1..3 only for
// r == 2).  Slots live on the (ny+1) x (nx+1) lattice; invalid combinations return -1.
WV_HD int64_t entity_dof(const Mesh &m, int i, int j, int kind) {
    switch (kind) {
    case 0: return dof_V(m, i, j);
    case 1: return (m.r == 2 && i < m.nx) ? dof_B(m, i, j) : -1;
    case 2: return (m.r == 2 && j < m.ny) ? dof_L(m, i, j) : -1;
    default: return (m.r == 2 && i < m.nx && j < m.ny) ? dof_D(m, i, j) : -1;
    }
}
WV_HD int64_t entity_dof_internal(const Mesh &m, int i, int j, int kind) {
    switch (kind) {
    case 0: return idof_V(m, i, j);
    case 1: return (m.r == 2 && i < m.nx) ? idof_B(m, i, j) : -1;
    case 2: return (m.r == 2 && j < m.ny) ? idof_L(m, i, j) : -1;
    default: return (m.r == 2 && i < m.nx && j < m.ny) ? idof_D(m, i, j) : -1;
    }
}
// support point of an entity (vertex, or edge midpoint as the mean of its two vertices)
WV_HD void entity_point(const Mesh &m, int i, int j, int kind, double &x, double &y) {
    const double xa = m.x0 + i * m.dx, ya = m.y0 + j * m.dy;
    if (kind == 0) { x = xa; y = ya; return; }
    const double xb = m.x0 + (i + 1) * m.dx, yb = m.y0 + (j + 1) * m.dy;
    if (kind == 1) { x = 0.5 * (xa + xb); y = ya; }
    else if (kind == 2) { x = xa; y = 0.5 * (ya + yb); }
    else { x = 0.5 * (xb + xa); y = 0.5 * (ya + yb); }
}
WV_HD bool entity_on_boundary(const Mesh &m, int i, int j, int kind) {
    switch (kind) {
    case 0: return i == 0 || i == m.nx || j == 0 || j == m.ny;
    case 1: return j == 0 || j == m.ny;
    case 2: return i == 0 || i == m.nx;
    default: return false;
    }
}
// cells adjacent to an entity (at most 6); returns the count
WV_HD int entity_cells(const Mesh &m, int i, int j, int kind, int64_t *cells) {
    int n = 0;
    auto quad_ok = [&](int qi, int qj) { return qi >= 0 && qi < m.nx && qj >= 0 && qj < m.ny; };
    auto cid = [&](int qi, int qj, int t) { return 2 * ((int64_t)qj * m.nx + qi) + t; };
    if (kind == 0) {
        if (quad_ok(i - 1, j - 1)) cells[n++] = cid(i - 1, j - 1, 1);
        if (quad_ok(i, j - 1)) { cells[n++] = cid(i, j - 1, 0); cells[n++] = cid(i, j - 1, 1); }
        if (quad_ok(i - 1, j)) { cells[n++] = cid(i - 1, j, 0); cells[n++] = cid(i - 1, j, 1); }
        if (quad_ok(i, j)) cells[n++] = cid(i, j, 0);
    } else if (kind == 1) {
        if (quad_ok(i, j - 1)) cells[n++] = cid(i, j - 1, 1);
        if (quad_ok(i, j)) cells[n++] = cid(i, j, 0);
    } else if (kind == 2) {
        if (quad_ok(i - 1, j)) cells[n++] = cid(i - 1, j, 1);
        if (quad_ok(i, j)) cells[n++] = cid(i, j, 0);
    } else {
        cells[n++] = cid(i, j, 0);
        cells[n++] = cid(i, j, 1);
    }
    return n;
}

// ---- FE_SimplexP(r) on the reference triangle (0,0),(1,0),(0,1) ---------------------------
WV_HD void shape_values(int r, double xi, double eta, double *phi) {
    const double l0 = 1.0 - xi - eta, l1 = xi, l2 = eta;
    if (r == 1) { phi[0] = l0; phi[1] = l1; phi[2] = l2; return; }
    phi[0] = l0 * (2 * l0 - 1); phi[1] = l1 * (2 * l1 - 1); phi[2] = l2 * (2 * l2 - 1);
    phi[3] = 4 * l0 * l1; phi[4] = 4 * l1 * l2; phi[5] = 4 * l2 * l0;
}
WV_HD void shape_grads(int r, double xi, double eta, double *dxi, double *deta) {
    const double l0 = 1.0 - xi - eta, l1 = xi, l2 = eta;
    if (r == 1) {
        dxi[0] = -1; dxi[1] = 1; dxi[2] = 0;
        deta[0] = -1; deta[1] = 0; deta[2] = 1;
        return;
    }
    dxi[0] = -(4 * l0 - 1); dxi[1] = 4 * l1 - 1; dxi[2] = 0;
    deta[0] = -(4 * l0 - 1); deta[1] = 0; deta[2] = 4 * l2 - 1;
    dxi[3] = 4 * (l0 - l1); deta[3] = -4 * l1;
    dxi[4] = 4 * l2; deta[4] = 4 * l1;
    dxi[5] = -4 * l2; deta[5] = 4 * (l0 - l2);
}

// ---- quadrature ([deal.II] QGaussSimplex<2>(n), weights sum to 1/2) -----------------------
// deal.II >= 9.4 tables: n=2: 4 points, degree 3 (Hillion).  n=3: 7 points, degree 5
// (Hammer-Marlowe-Stroud).  n=4: 15 points, degree 7 (Witherden-Vincent).  Filled on the host,
// see quadrature.cpp.
struct Quadrature {
    int nq;
    double xi[16], eta[16], w[16];
};

}  // namespace wv
