// ORACLE (test infrastructure, NOT product code).
//
// Exact sign-of-determinant arithmetic for the CPU restatement of the un-vendored
// third-party library qnzhou/simplicial_arrangement (GIT_TAG main, un-pinned;
// /root/reference/cmake/simplicial_arrangement.cmake:5-10).  Upstream gets its
// exact signs from an "implicit/indirect predicates" package; the only property the
// reference relies on is that the SIGN is exact, so any exact method yields identical
// combinatorics.  This file uses Shewchuk-style floating-point expansions held in
// std::vector<double> (deliberately different in form from the fixed-capacity
// device implementation in robust-implicit-surface-networks_b200/csrc/predicates.cuh).
//
// Assumptions: IEEE-754 binary64, round-to-nearest-even, no FMA contraction
// (-ffp-contract=off), no overflow/underflow in the products (inputs are O(1e-300..1e75)).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace sa_oracle {

using Expansion = std::vector<double>; // non-overlapping, increasing magnitude, no zeros

inline void two_sum(double a, double b, double& x, double& y)
{
    x = a + b;
    double bv = x - a;
    double av = x - bv;
    double br = b - bv;
    double ar = a - av;
    y = ar + br;
}

inline void two_prod(double a, double b, double& x, double& y)
{
    x = a * b;
    y = std::fma(a, b, -x); // exact rounding error of the product
}

// e + b  (Shewchuk GROW-EXPANSION with zero elimination)
inline Expansion grow(const Expansion& e, double b)
{
    Expansion h;
    h.reserve(e.size() + 1);
    double q = b;
    for (double ei : e) {
        double qn, lo;
        two_sum(q, ei, qn, lo);
        if (lo != 0.0) h.push_back(lo);
        q = qn;
    }
    if (q != 0.0) h.push_back(q);
    return h;
}

// e + f (EXPANSION-SUM)
inline Expansion sum(const Expansion& e, const Expansion& f)
{
    Expansion h = e;
    for (double fi : f) h = grow(h, fi);
    return h;
}

inline Expansion neg(const Expansion& e)
{
    Expansion h = e;
    for (double& x : h) x = -x;
    return h;
}

// e * b (SCALE-EXPANSION with zero elimination)
inline Expansion scale(const Expansion& e, double b)
{
    Expansion h;
    if (e.empty() || b == 0.0) return h;
    h.reserve(2 * e.size());
    double q, lo;
    two_prod(e[0], b, q, lo);
    if (lo != 0.0) h.push_back(lo);
    for (size_t i = 1; i < e.size(); ++i) {
        double t_hi, t_lo;
        two_prod(e[i], b, t_hi, t_lo);
        double qq, l2;
        two_sum(q, t_lo, qq, l2);
        if (l2 != 0.0) h.push_back(l2);
        // FAST-TWO-SUM(t_hi, qq)
        double qn = t_hi + qq;
        double l3 = qq - (qn - t_hi);
        if (l3 != 0.0) h.push_back(l3);
        q = qn;
    }
    if (q != 0.0) h.push_back(q);
    return h;
}

inline Expansion mul(const Expansion& e, const Expansion& f)
{
    Expansion h;
    for (double fi : f) h = sum(h, scale(e, fi));
    return h;
}

inline Expansion from_double(double a)
{
    Expansion e;
    if (a != 0.0) e.push_back(a);
    return e;
}

// exact a - b as an expansion
inline Expansion diff(double a, double b)
{
    double x, y;
    two_sum(a, -b, x, y);
    Expansion e;
    if (y != 0.0) e.push_back(y);
    if (x != 0.0) e.push_back(x);
    return e;
}

inline int sign(const Expansion& e)
{
    if (e.empty()) return 0;
    return e.back() > 0 ? 1 : -1;
}

// Exact determinant of an n x n matrix of expansions (Laplace expansion along row 0).
inline Expansion det_exact(int n, const std::vector<std::vector<Expansion>>& m)
{
    if (n == 1) return m[0][0];
    if (n == 2) return sum(mul(m[0][0], m[1][1]), neg(mul(m[0][1], m[1][0])));
    Expansion acc;
    for (int c = 0; c < n; ++c) {
        if (m[0][c].empty()) continue;
        std::vector<std::vector<Expansion>> sub(n - 1, std::vector<Expansion>(n - 1));
        for (int r = 1; r < n; ++r) {
            int cc = 0;
            for (int k = 0; k < n; ++k) {
                if (k == c) continue;
                sub[r - 1][cc++] = m[r][k];
            }
        }
        Expansion term = mul(m[0][c], det_exact(n - 1, sub));
        acc = sum(acc, (c & 1) ? neg(term) : term);
    }
    return acc;
}

// Floating-point determinant together with its "permanent" (same formula on absolute
// values); |det_fp - det_true| <= c_n * eps * perm for a small constant c_n.
inline void det_fp(int n, const double* m /*row-major n*n*/, double& det, double& perm)
{
    if (n == 1) {
        det = m[0];
        perm = std::fabs(m[0]);
        return;
    }
    if (n == 2) {
        double a = m[0] * m[3], b = m[1] * m[2];
        det = a - b;
        perm = std::fabs(a) + std::fabs(b);
        return;
    }
    det = 0;
    perm = 0;
    double sub[9];
    for (int c = 0; c < n; ++c) {
        for (int r = 1; r < n; ++r) {
            int cc = 0;
            for (int k = 0; k < n; ++k) {
                if (k == c) continue;
                sub[(r - 1) * (n - 1) + cc++] = m[r * n + k];
            }
        }
        double d, p;
        det_fp(n - 1, sub, d, p);
        double t = m[c] * d;
        det = (c & 1) ? det - t : det + t;
        perm += std::fabs(m[c]) * p;
    }
}

// Sign of det of an n x n matrix of doubles: semi-static filter, exact fallback.
// `n_exact` (optional) counts how often the fallback ran.
inline int det_sign(int n, const double* m, size_t* n_exact = nullptr)
{
    double det, perm;
    det_fp(n, m, det, perm);
    const double eps = 1.1102230246251565e-16; // 2^-53
    const double bound = 64.0 * eps * perm;
    if (perm > 1e-280 && perm < 1e280) {
        if (det > bound) return 1;
        if (det < -bound) return -1;
    }
    if (n_exact) ++*n_exact;
    std::vector<std::vector<Expansion>> e(n, std::vector<Expansion>(n));
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) e[r][c] = from_double(m[r * n + c]);
    return sign(det_exact(n, e));
}

// Sign of det of an n x n matrix whose entries are exact differences a[r][c]-b[r][c]
// (used by the material-interface predicates; entry (r,c) = a - b, rows flagged `plain`
// use a alone).  Filtered with an error bound that also covers the rounding of the
// differences; exact fallback on 2-term expansions.
inline int det_sign_diff(int n, const double* a, const double* b, size_t* n_exact = nullptr)
{
    double m[16];
    for (int i = 0; i < n * n; ++i) m[i] = a[i] - b[i];
    double det, perm;
    det_fp(n, m, det, perm);
    const double eps = 1.1102230246251565e-16;
    const double bound = 128.0 * eps * perm;
    if (perm > 1e-280 && perm < 1e280) {
        if (det > bound) return 1;
        if (det < -bound) return -1;
    }
    if (n_exact) ++*n_exact;
    std::vector<std::vector<Expansion>> e(n, std::vector<Expansion>(n));
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) e[r][c] = diff(a[r * n + c], b[r * n + c]);
    return sign(det_exact(n, e));
}

} // namespace sa_oracle
