/*
 * gate_matrix.cpp — the 16 single-qubit matrices of qgate's gate set, in double.
 *
 * Values (including qgate's global-phase conventions for U and U2, which differ from
 * OpenQASM by exp(i (lambda + phi) / 2)) follow qgate/simulator/src/GateMatrix.cpp:13-158
 * and qgate/simulator/pymatrix.py:18-149; the adjoint is the conjugate transpose
 * (GateMatrix.cpp:160-168).  tests/test_gate_matrix.py checks every entry against the
 * golden matrices recorded from the reference.
 */
#include <cmath>

#include "gate_matrix.h"

namespace qgb {

namespace {

struct C {
    double re, im;
};
inline C expi(double x) { return C{std::cos(x), std::sin(x)}; } /* e^{ix} */
inline C scale(C a, double s) { return C{a.re * s, a.im * s}; }
inline C neg(C a) { return C{-a.re, -a.im}; }

inline void put(double *m, int idx, C v) {
    m[2 * idx] = v.re;
    m[2 * idx + 1] = v.im;
}

const int kNumArgs[16] = {3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0};

} // namespace

int gate_matrix_n_args(int gate_id) { return (gate_id < 0 || gate_id >= 16) ? -1 : kNumArgs[gate_id]; }

bool gate_matrix(int gate_id, const double *a, int adjoint, double *m) {
    const C one{1., 0.}, zero{0., 0.}, I{0., 1.};
    const double r = std::sqrt(0.5);
    C m00 = zero, m01 = zero, m10 = zero, m11 = zero;
    switch (gate_id) {
    case 0: { /* U(theta, phi, lambda) */
        const double c = std::cos(a[0] * 0.5), s = std::sin(a[0] * 0.5);
        const double p = a[1] * 0.5, l = a[2] * 0.5;
        m00 = scale(expi(-l - p), c);
        m01 = neg(scale(expi(l - p), s));
        m10 = scale(expi(-l + p), s);
        m11 = scale(expi(l + p), c);
        break;
    }
    case 1: { /* U2(phi, lambda) */
        const double p = a[0] * 0.5, l = a[1] * 0.5;
        m00 = scale(expi(-l - p), r);
        m01 = neg(scale(expi(l - p), r));
        m10 = scale(expi(-l + p), r);
        m11 = scale(expi(l + p), r);
        break;
    }
    case 2: /* U1(lambda) */
        m00 = one;
        m11 = expi(a[0]);
        break;
    case 3: /* ID */
        m00 = one;
        m11 = one;
        break;
    case 4: /* X */
        m01 = one;
        m10 = one;
        break;
    case 5: /* Y */
        m01 = neg(I);
        m10 = I;
        break;
    case 6: /* Z */
        m00 = one;
        m11 = neg(one);
        break;
    case 7: /* H */
        m00 = m01 = m10 = C{r, 0.};
        m11 = C{-r, 0.};
        break;
    case 8: /* S */
        m00 = one;
        m11 = I;
        break;
    case 9: /* T */
        m00 = one;
        m11 = expi(M_PI * 0.25);
        break;
    case 10: { /* RX(theta) */
        const double c = std::cos(a[0] / 2.), s = std::sin(a[0] / 2.);
        m00 = m11 = C{c, 0.};
        m01 = m10 = C{0., -s};
        break;
    }
    case 11: { /* RY(theta) */
        const double c = std::cos(a[0] / 2.), s = std::sin(a[0] / 2.);
        m00 = m11 = C{c, 0.};
        m01 = C{-s, 0.};
        m10 = C{s, 0.};
        break;
    }
    case 12: /* RZ(theta) */
        m00 = expi(-(a[0] / 2.));
        m11 = expi(a[0] / 2.);
        break;
    case 13: /* ExpiI(theta) = e^{i theta} I */
        m00 = m11 = expi(a[0]);
        break;
    case 14: /* ExpiZ(theta) = diag(e^{i theta}, e^{-i theta}) */
        m00 = expi(a[0]);
        m11 = expi(-a[0]);
        break;
    case 15: /* SH = S H */
        m00 = m01 = C{r, 0.};
        m10 = C{0., r};
        m11 = C{0., -r};
        break;
    default:
        return false;
    }
    if (adjoint) {
        const C t = m01;
        m00 = C{m00.re, -m00.im};
        m01 = C{m10.re, -m10.im};
        m10 = C{t.re, -t.im};
        m11 = C{m11.re, -m11.im};
    }
    put(m, 0, m00);
    put(m, 1, m01);
    put(m, 2, m10);
    put(m, 3, m11);
    return true;
}

} // namespace qgb
