// quadrature.cpp -- simplex quadrature tables of the product (host side; copied to the device
// as plain data).  [deal.II] QGaussSimplex<2>(n) at the reference's call sites
// src/WaveEquationBase.cpp:82 (n = r+1, assembly) and :371,405 (n = r+2, error norms).
#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "mesh.h"

namespace wv {
namespace {

// Legendre P_n(x) and P_n'(x)
void legendre(int n, double x, double &p, double &dp) {
    double p0 = 1.0, p1 = x;
    if (n == 0) { p = 1.0; dp = 0.0; return; }
    for (int k = 2; k <= n; ++k) {
        const double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        p0 = p1; p1 = pk;
    }
    p = p1;
    dp = n * (x * p1 - p0) / (x * x - 1.0);
}

void gauss_legendre(int n, double *x, double *w) {
    for (int k = 0; k < n; ++k) {
        double z = -std::cos(M_PI * (k + 0.75) / (n + 0.5));
        for (int it = 0; it < 100; ++it) {
            double p, dp;
            legendre(n, z, p, dp);
            const double dz = p / dp;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        double p, dp;
        legendre(n, z, p, dp);
        x[k] = z;
        w[k] = 2.0 / ((1.0 - z * z) * dp * dp);
    }
}

// n-point Gauss rule for the weight (1-x) on [-1,1]: nodes are the zeros of
// (P_n(x) - P_{n+1}(x)) / (1 - x); weights from the moment equations.
void gauss_jacobi10(int n, double *x, double *w) {
    auto f = [&](double z, double &v, double &dv) {
        double pa, da, pb, db;
        legendre(n, z, pa, da);
        legendre(n + 1, z, pb, db);
        v = pa - pb;
        dv = da - db;
    };
    for (int k = 0; k < n; ++k) {
        double z = -std::cos(M_PI * (k + 0.5) / (n + 0.5));
        for (int it = 0; it < 200; ++it) {
            double v, dv, s = 1.0 / (z - 1.0);  // deflate the root at x = 1 and the found ones
            f(z, v, dv);
            for (int m = 0; m < k; ++m) s += 1.0 / (z - x[m]);
            const double dz = v / (dv - v * s);
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        x[k] = z;
    }
    std::sort(x, x + n);
    // moments  int_{-1}^{1} (1-x) x^m dx
    double A[8][9];
    for (int m = 0; m < n; ++m) {
        for (int k = 0; k < n; ++k) A[m][k] = std::pow(x[k], m);
        const double a = (m % 2 == 0) ? 2.0 / (m + 1) : 0.0;        // int x^m
        const double b = ((m + 1) % 2 == 0) ? 2.0 / (m + 2) : 0.0;  // int x^(m+1)
        A[m][n] = a - b;
    }
    for (int c = 0; c < n; ++c) {  // Gaussian elimination with partial pivoting
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        for (int k = 0; k <= n; ++k) std::swap(A[c][k], A[piv][k]);
        for (int r = c + 1; r < n; ++r) {
            const double fct = A[r][c] / A[c][c];
            for (int k = c; k <= n; ++k) A[r][k] -= fct * A[c][k];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        double s = A[r][n];
        for (int k = r + 1; k < n; ++k) s -= A[r][k] * w[k];
        w[r] = s / A[r][r];
    }
}

}  // namespace

Quadrature make_quadrature(int n1d) {
    Quadrature q{};
    if (n1d == 2) {
        q.nq = 3;
        const double pts[3][2] = {{1.0 / 6, 1.0 / 6}, {2.0 / 3, 1.0 / 6}, {1.0 / 6, 2.0 / 3}};
        for (int k = 0; k < 3; ++k) { q.xi[k] = pts[k][0]; q.eta[k] = pts[k][1]; q.w[k] = 1.0 / 6; }
    } else if (n1d == 3) {
        q.nq = 7;
        const double s15 = std::sqrt(15.0);
        const double a = (6.0 - s15) / 21.0, b = (6.0 + s15) / 21.0;
        const double wa = (155.0 - s15) / 2400.0, wb = (155.0 + s15) / 2400.0;
        const double pts[7][3] = {{1.0 / 3, 1.0 / 3, 9.0 / 80}, {a, a, wa}, {1 - 2 * a, a, wa},
                                  {a, 1 - 2 * a, wa},           {b, b, wb}, {1 - 2 * b, b, wb},
                                  {b, 1 - 2 * b, wb}};
        for (int k = 0; k < 7; ++k) { q.xi[k] = pts[k][0]; q.eta[k] = pts[k][1]; q.w[k] = pts[k][2]; }
    } else if (n1d == 4) {
        const int n = 4;
        double gx[8], gw[8], jx[8], jw[8];
        gauss_legendre(n, gx, gw);
        gauss_jacobi10(n, jx, jw);
        q.nq = n * n;
        int k = 0;
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                q.xi[k] = 0.5 * (1.0 + jx[a]);
                q.eta[k] = 0.25 * (1.0 - jx[a]) * (1.0 + gx[b]);
                q.w[k] = jw[a] * gw[b] / 8.0;
                ++k;
            }
    } else
        throw std::invalid_argument("unsupported quadrature order");
    return q;
}

}  // namespace wv
