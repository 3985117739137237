// quadrature.cpp -- simplex quadrature tables of the product (host side; copied to the device
// as plain data).  [deal.II] QGaussSimplex<2>(n) at the reference's call sites
// src/WaveEquationBase.cpp:82 (n = r+1, assembly and forcing) and :371,405 (n = r+2, error norms).
//
// The tables are those of deal.II >= 9.4 (9.3 has no n = 4 rule in two dimensions, and the
// reference's own P2 error tables, analysis/data/convergence-results.csv, are reproduced to their
// 7 printed digits only by the 15-point Witherden-Vincent rule):
//   n = 2   4 points, degree 3   Hillion's scheme: 2x2 Gauss x Gauss-Jacobi(1,0), collapsed
//   n = 3   7 points, degree 5   Hammer-Marlowe-Stroud (Radon)
//   n = 4  15 points, degree 7   Witherden-Vincent
// Weights sum to 1/2.  The n = 2 rule is not invariant under vertex permutations, so the cell's
// vertex order (mesh.h cell_geometry: xi along v0->v1, eta along v0->v2) is part of the result
// whenever the integrand is not a cubic (forcing vectors, variable c).
#include <algorithm>
#include <array>
#include <cmath>
#include <stdexcept>

#include "mesh.h"

namespace wv {
namespace {

struct Builder {
    Quadrature q{};
    void add(double xi, double eta, double w) {
        if (q.nq >= 16) throw std::logic_error("quadrature table overflow");
        q.xi[q.nq] = xi;
        q.eta[q.nq] = eta;
        q.w[q.nq] = w;
        ++q.nq;
    }
    // every distinct arrangement of the barycentric triple, ascending lexicographic order; the
    // Cartesian point is the first two coordinates
    void add_orbit(std::array<double, 3> b, double w) {
        std::sort(b.begin(), b.end());
        do add(b[0], b[1], w);
        while (std::next_permutation(b.begin(), b.end()));
    }
};

Quadrature hillion4() {
    Builder B;
    // eta: Gauss-Jacobi(1,0) nodes (4 -+ sqrt 6)/10 on [0,1]; xi = (1 - eta)(1 -+ 1/sqrt 3)/2
    const double wl = 0.31804138174397717, wh = 0.18195861825602283;  // (9 +- sqrt 6)/36
    B.add(0.17855872826361643, 0.1550510257216822, 0.5 * wl);
    B.add(0.07503111022260812, 0.6449489742783178, 0.5 * wh);
    B.add(0.6663902460147014, 0.1550510257216822, 0.5 * wl);
    B.add(0.28001991549907407, 0.6449489742783178, 0.5 * wh);
    return B.q;
}

Quadrature hammer_marlowe_stroud7() {
    Builder B;
    const double r = std::sqrt(15.0);
    const double lo = 2.0 / 7.0 - r / 21.0, hi = 2.0 / 7.0 + r / 21.0;
    const double lo_c = 3.0 / 7.0 + 2.0 * r / 21.0, hi_c = 3.0 / 7.0 - 2.0 * r / 21.0;  // 1 - 2 lo, 1 - 2 hi
    const double w_lo = 0.5 * (31.0 / 240.0 - r / 1200.0), w_hi = 0.5 * (31.0 / 240.0 + r / 1200.0);
    B.add(1.0 / 3.0, 1.0 / 3.0, 0.5 * (9.0 / 40.0));
    B.add(lo_c, lo, w_lo);
    B.add(lo, lo_c, w_lo);
    B.add(lo, lo, w_lo);
    B.add(hi_c, hi, w_hi);
    B.add(hi, hi_c, w_hi);
    B.add(hi, hi, w_hi);
    return B.q;
}

Quadrature witherden_vincent15() {
    Builder B;
    struct S21 { double a, w; };
    const S21 s21[3] = {{0.03373064855458785, 0.016545050110792132},
                        {0.24157738259540357, 0.12794417123015558},
                        {0.47430969250471822, 0.07708664618598607}};
    for (const S21 &o : s21) B.add_orbit({o.a, o.a, 1.0 - 2.0 * o.a}, 0.5 * o.w);
    const double b1 = 0.047036644652595234, b2 = 0.19868331479735159;
    B.add_orbit({b1, b2, 1.0 - b1 - b2}, 0.5 * 0.05587873290319978);
    return B.q;
}

}  // namespace

Quadrature make_quadrature(int n1d) {
    switch (n1d) {
        case 2: return hillion4();
        case 3: return hammer_marlowe_stroud7();
        case 4: return witherden_vincent15();
        default: throw std::invalid_argument("unsupported quadrature order");
    }
}

}  // namespace wv
