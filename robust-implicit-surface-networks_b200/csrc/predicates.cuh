// Exact sign predicates for the per-tet kernels (sm_100a).
//
// Replaces the exact arithmetic of the un-vendored qnzhou/simplicial_arrangement library that
// the reference calls at /root/reference/src/implicit_arrangement.cpp:279,283 and
// src/material_interface.cpp:323,327.  Every predicate is a sign of a small determinant of the
// raw function values: a floating-point evaluation with a semi-static error bound decides the
// common case; otherwise the determinant is evaluated exactly with fixed-capacity floating-point
// expansions (two_sum / two_prod via fma, grow / scale with zero elimination) held in local
// memory.  The file must be compiled with -fmad=false so that no a*b+c is contracted.
#pragma once
#include <cstdint>

namespace rin {

#define RIN_EPS 1.1102230246251565e-16 /* 2^-53 */

__device__ __forceinline__ void two_sum(double a, double b, double& x, double& y)
{
    x = __dadd_rn(a, b);
    double bv = __dsub_rn(x, a);
    double av = __dsub_rn(x, bv);
    double br = __dsub_rn(b, bv);
    double ar = __dsub_rn(a, av);
    y = __dadd_rn(ar, br);
}

__device__ __forceinline__ void two_prod(double a, double b, double& x, double& y)
{
    x = __dmul_rn(a, b);
    y = __fma_rn(a, b, -x);
}

// h = e + b, in place (h may alias e); returns the new length
__device__ __forceinline__ int exp_grow(double* e, int n, double b)
{
    double q = b;
    int m = 0;
    for (int i = 0; i < n; ++i) {
        double qn, lo;
        two_sum(q, e[i], qn, lo);
        if (lo != 0.0) e[m++] = lo;
        q = qn;
    }
    if (q != 0.0) e[m++] = q;
    return m;
}

// acc += sgn * f   (sgn = +1 / -1)
__device__ __forceinline__ int exp_add(double* acc, int n, const double* f, int nf, double sgn)
{
    for (int i = 0; i < nf; ++i) n = exp_grow(acc, n, sgn * f[i]);
    return n;
}

// h = e * b (h must not alias e, capacity 2*n); returns the length
__device__ __forceinline__ int exp_scale(const double* e, int n, double b, double* h)
{
    if (n == 0 || b == 0.0) return 0;
    int m = 0;
    double q, lo;
    two_prod(e[0], b, q, lo);
    if (lo != 0.0) h[m++] = lo;
    for (int i = 1; i < n; ++i) {
        double th, tl;
        two_prod(e[i], b, th, tl);
        double qq, l2;
        two_sum(q, tl, qq, l2);
        if (l2 != 0.0) h[m++] = l2;
        double qn = __dadd_rn(th, qq);
        double l3 = __dsub_rn(qq, __dsub_rn(qn, th));
        if (l3 != 0.0) h[m++] = l3;
        q = qn;
    }
    if (q != 0.0) h[m++] = q;
    return m;
}

__device__ __forceinline__ int exp_sign(const double* e, int n)
{
    return n == 0 ? 0 : (e[n - 1] > 0 ? 1 : -1);
}

// ---- exact determinants of doubles (row-major) ------------------------------------------------
// 2x2: <= 4 components
__device__ __noinline__ int det2_exact(double a, double b, double c, double d, double* out)
{
    double x, y;
    int n = 0;
    two_prod(a, d, x, y);
    if (y != 0.0) out[n++] = y;
    if (x != 0.0) out[n++] = x;
    two_prod(b, c, x, y);
    n = exp_grow(out, n, -y);
    n = exp_grow(out, n, -x);
    return n;
}

// 3x3: <= 24 components
__device__ __noinline__ int det3_exact(const double* m, double* out)
{
    double m2[4], sc[8];
    int n = 0;
    for (int c = 0; c < 3; ++c) {
        if (m[c] == 0.0) continue;
        int c0 = (c == 0) ? 1 : 0, c1 = (c == 2) ? 1 : 2;
        int k = det2_exact(m[3 + c0], m[3 + c1], m[6 + c0], m[6 + c1], m2);
        int ks = exp_scale(m2, k, m[c], sc);
        n = exp_add(out, n, sc, ks, (c & 1) ? -1.0 : 1.0);
    }
    return n;
}

// 4x4: <= 192 components
__device__ __noinline__ int det4_exact(const double* m, double* out)
{
    double sub[9], d3[24], sc[48];
    int n = 0;
    for (int c = 0; c < 4; ++c) {
        if (m[c] == 0.0) continue;
        for (int r = 1; r < 4; ++r) {
            int cc = 0;
            for (int k = 0; k < 4; ++k)
                if (k != c) sub[(r - 1) * 3 + cc++] = m[r * 4 + k];
        }
        int k3 = det3_exact(sub, d3);
        int ks = exp_scale(d3, k3, m[c], sc);
        n = exp_add(out, n, sc, ks, (c & 1) ? -1.0 : 1.0);
    }
    return n;
}

// ---- filtered signs ---------------------------------------------------------------------------
// Each returns the exact sign; *n_exact is incremented when the expansion fallback ran.

__device__ __forceinline__ bool filter_ok(double det, double perm, double c, int& s)
{
    double bound = c * RIN_EPS * perm;
    if (perm > 1e-280 && perm < 1e280) {
        if (det > bound) {
            s = 1;
            return true;
        }
        if (det < -bound) {
            s = -1;
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ int det2_sign(double a, double b, double c, double d, unsigned* n_exact)
{
    double p = a * d, q = b * c;
    int s;
    if (filter_ok(p - q, fabs(p) + fabs(q), 8.0, s)) return s;
    ++*n_exact;
    double e[4];
    return exp_sign(e, det2_exact(a, b, c, d, e));
}

__device__ __forceinline__ void det3_fp(const double* m, double& det, double& perm)
{
    double a = m[4] * m[8], b = m[5] * m[7];
    double c = m[3] * m[8], d = m[5] * m[6];
    double e = m[3] * m[7], f = m[4] * m[6];
    det = m[0] * (a - b) - m[1] * (c - d) + m[2] * (e - f);
    perm = fabs(m[0]) * (fabs(a) + fabs(b)) + fabs(m[1]) * (fabs(c) + fabs(d)) +
           fabs(m[2]) * (fabs(e) + fabs(f));
}

__device__ __forceinline__ int det3_sign(const double* m, unsigned* n_exact)
{
    double det, perm;
    det3_fp(m, det, perm);
    int s;
    if (filter_ok(det, perm, 32.0, s)) return s;
    ++*n_exact;
    double e[24];
    return exp_sign(e, det3_exact(m, e));
}

__device__ __forceinline__ int det4_sign(const double* m, unsigned* n_exact)
{
    double det = 0, perm = 0;
    double sub[9];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        for (int r = 1; r < 4; ++r) {
            int cc = 0;
            for (int k = 0; k < 4; ++k)
                if (k != c) sub[(r - 1) * 3 + cc++] = m[r * 4 + k];
        }
        double d, p;
        det3_fp(sub, d, p);
        double t = m[c] * d;
        det = (c & 1) ? det - t : det + t;
        perm += fabs(m[c]) * p;
    }
    int s;
    if (filter_ok(det, perm, 64.0, s)) return s;
    ++*n_exact;
    double e[192];
    return exp_sign(e, det4_exact(m, e));
}

__device__ __forceinline__ int detn_sign(int n, const double* m, unsigned* n_exact)
{
    if (n == 2) return det2_sign(m[0], m[1], m[2], m[3], n_exact);
    if (n == 3) return det3_sign(m, n_exact);
    return det4_sign(m, n_exact);
}

} // namespace rin
