// General per-tet material interface (>= 3 materials, ties, or lookup disabled).
//
// Replaces compute_material_interface() of the un-vendored qnzhou/simplicial_arrangement library,
// called by the reference at /root/reference/src/material_interface.cpp:323,327.  One thread owns
// one tetrahedron.  Materials are inserted one at a time; every cell carries the material that is
// maximal in it; inserting M classifies every vertex by the exact sign of (M - current maximum),
// cuts every cell by M = its material and merges all positive parts into M's single convex cell,
// dropping what lies strictly inside it and merging the simplex-boundary pieces.  Conventions
// (ordering, orientation, labels) are those of DESIGN.md "per-tet complex conventions" so that
// local ids agree with the CPU oracle bit for bit.
#pragma once
#include "ia_complex.cuh"

namespace rin {

struct MICaps
{
    using idx = uint16_t;
    static constexpr int MAXK = 64;
    static constexpr int MAXV = 200;
    static constexpr int MAXE = 480;
    static constexpr int MAXF = 320;
    static constexpr int MAXC = 96;
    static constexpr int MAXFE = 1600;
    static constexpr int MAXLOOP = 48;
};
struct MICapsSmall
{
    using idx = uint8_t;
    static constexpr int MAXK = 8;
    static constexpr int MAXV = 44;
    static constexpr int MAXE = 96;
    static constexpr int MAXF = 64;
    static constexpr int MAXC = 16;
    static constexpr int MAXFE = 320;
    static constexpr int MAXLOOP = 16;
};

// sign of det of an n x n matrix whose entries are a[i] - b[i] (exact: every entry is expanded
// into the two-term exact difference in the fallback)
__device__ __noinline__ int detn_diff_exact_sign(int n, const double* a, const double* b)
{
    // multilinear expansion over the choice (a or -b) per row would need 2^n determinants of
    // doubles; instead expand every entry exactly as hi+lo = two_sum(a,-b) and use linearity per
    // row: det = sum over subsets of rows taking lo.  n <= 4 -> at most 16 determinants.
    double hi[16], lo[16];
    for (int i = 0; i < n * n; ++i) two_sum(a[i], -b[i], hi[i], lo[i]);
    double acc[192 * 2];
    int na = 0;
    double m[16], e[192];
    for (int mask = 0; mask < (1 << n); ++mask) {
        bool zero_row = false;
        for (int r = 0; r < n; ++r) {
            bool all0 = true;
            for (int c = 0; c < n; ++c) {
                double x = ((mask >> r) & 1) ? lo[r * n + c] : hi[r * n + c];
                m[r * n + c] = x;
                all0 &= (x == 0.0);
            }
            zero_row |= all0;
        }
        if (zero_row) continue;
        int ne = 0;
        if (n == 2)
            ne = det2_exact(m[0], m[1], m[2], m[3], e);
        else if (n == 3)
            ne = det3_exact(m, e);
        else
            ne = det4_exact(m, e);
        na = exp_add(acc, na, e, ne, 1.0);
        if (na > 192 * 2 - 200) { // compress: never reached for non-overlapping inputs of this size
        }
    }
    return exp_sign(acc, na);
}

__device__ __forceinline__ int detn_diff_sign(int n, const double* a, const double* b, unsigned* n_exact)
{
    double m[16];
    for (int i = 0; i < n * n; ++i) m[i] = a[i] - b[i];
    double det, perm;
    if (n == 2) {
        double p = m[0] * m[3], q = m[1] * m[2];
        det = p - q;
        perm = fabs(p) + fabs(q);
    } else if (n == 3) {
        det3_fp(m, det, perm);
    } else {
        det = 0;
        perm = 0;
        double sub[9];
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
    }
    int s;
    if (filter_ok(det, perm, 160.0, s)) return s; // bound covers the rounding of the differences
    ++*n_exact;
    return detn_diff_exact_sign(n, a, b);
}

template <class Caps>
struct MIComplex
{
    using I = typename Caps::idx;
    static constexpr I NI = (I)~(I)0;
    double mval[Caps::MAXK][4]; // input material j (id 4+j) at the 4 corners
    int nm;                     // next material id
    int nv, ne, nf, nc, nfe;
    uint8_t vm[Caps::MAXV][4];
    int8_t vo[Caps::MAXV];
    I vmap[Caps::MAXV];
    I ev0[Caps::MAXE], ev1[Caps::MAXE];
    uint8_t em[Caps::MAXE][3];
    I ec_pos[Caps::MAXE], ec_neg[Caps::MAXE], ec_x[Caps::MAXE]; // ec_pos reused as edge remap
    uint8_t ec_split[Caps::MAXE];
    I merged_of[Caps::MAXE];
    // faces, double buffered (rebuilt on every insertion)
    uint16_t foff[2][Caps::MAXF];
    uint8_t flen[2][Caps::MAXF], fb[2][Caps::MAXF], fpos[2][Caps::MAXF], fneg[2][Caps::MAXF];
    I fv[2][Caps::MAXFE], fe[2][Caps::MAXFE];
    I fc_pos[Caps::MAXF], fc_neg[Caps::MAXF], fc_cut[Caps::MAXF];
    uint8_t fc_split[Caps::MAXF];
    uint8_t cmat[Caps::MAXC], cstat[Caps::MAXC], cmap[Caps::MAXC], cneg[Caps::MAXC];
    uint8_t umi[Caps::MAXK + 4]; // material -> group (duplicates)
    int n_groups;
    bool has_dup;
    int cur; // active face buffer
    int err;
    unsigned n_exact;

    __device__ void init(const double v[4])
    {
        err = 0;
        n_exact = 0;
        cur = 0;
        nm = 5;
        for (int c = 0; c < 4; ++c) mval[0][c] = v[c];
        nv = 4;
        const uint8_t vms[4][4] = {{1, 2, 3, 4}, {0, 2, 3, 4}, {0, 1, 3, 4}, {0, 1, 2, 4}};
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k) vm[i][k] = vms[i][k];
        const uint8_t ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};
        const uint8_t epa[6] = {2, 1, 1, 0, 0, 0}, epb[6] = {3, 3, 2, 3, 2, 1};
        ne = 6;
        for (int e = 0; e < 6; ++e) {
            ev0[e] = ea[e];
            ev1[e] = eb[e];
            em[e][0] = epa[e];
            em[e][1] = epb[e];
            em[e][2] = 4;
        }
        const uint8_t eid[4][4] = {{0, 0, 1, 2}, {0, 0, 3, 4}, {1, 3, 0, 5}, {2, 4, 5, 0}};
        const uint8_t loops[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}}; // CCW from outside
        nf = 4;
        nfe = 0;
        for (int f = 0; f < 4; ++f) {
            foff[0][f] = nfe;
            flen[0][f] = 3;
            fb[0][f] = f;
            fpos[0][f] = N8;
            fneg[0][f] = 0;
            for (int k = 0; k < 3; ++k) {
                fv[0][nfe] = loops[f][k];
                fe[0][nfe] = eid[loops[f][k]][loops[f][(k + 1) % 3]];
                ++nfe;
            }
        }
        nc = 1;
        cmat[0] = 4;
        for (int p = 0; p < 5; ++p) umi[p] = p;
        n_groups = 5;
        has_dup = false;
    }

    // exact sign of (M - current maximum) at vertex v
    // nex / bad: exact-fallback counter and degeneracy flag of the caller (per lane in the warp version)
    __device__ int orient_vertex(int v, const double* M, unsigned& nex, int& bad)
    {
        const double* real[4];
        int k = 0;
        unsigned fixed = 0;
        for (int q = 0; q < 4; ++q) {
            int id = vm[v][q];
            if (id < 4)
                fixed |= 1u << id;
            else
                real[k++] = mval[id - 4];
        }
        int idx[4], n = 0;
        for (int c = 0; c < 4; ++c)
            if (!((fixed >> c) & 1)) idx[n++] = c;
        if (n == 1) {
            double a = M[idx[0]], b = real[0][idx[0]];
            return a > b ? 1 : (a < b ? -1 : 0);
        }
        double qa[16], qb[16], da[16], db[16];
        for (int r = 0; r < k - 1; ++r)
            for (int c = 0; c < n; ++c) {
                qa[r * n + c] = da[r * n + c] = real[r][idx[c]];
                qb[r * n + c] = db[r * n + c] = real[r + 1][idx[c]];
            }
        for (int c = 0; c < n; ++c) {
            qa[(k - 1) * n + c] = M[idx[c]];
            qb[(k - 1) * n + c] = real[0][idx[c]];
            da[(k - 1) * n + c] = 1.0;
            db[(k - 1) * n + c] = 0.0;
        }
        int sq = detn_diff_sign(n, qa, qb, &nex);
        if (sq == 0) return 0;
        int sd = detn_diff_sign(n, da, db, &nex);
        if (sd == 0) bad = 2;
        return sq * sd;
    }

    __device__ bool edge_is_positive(int e, int nE) const
    {
        if (e < nE) return !ec_split[e] && ec_pos[e] == e;
        return vo[ev0[e]] > 0 || vo[ev1[e]] > 0;
    }

    // inserts material `mid` (values in mval[mid-4]); returns a material it duplicates, or -1
    __device__ int add_material(int mid)
    {
        const double* M = mval[mid - 4];
        const int B = cur, B2 = cur ^ 1;
        for (int v = 0; v < nv; ++v) vo[v] = (int8_t)orient_vertex(v, M, n_exact, err);
        if (err) return -1;
        // ---- edges
        const int nE = ne;
        for (int e = 0; e < nE; ++e) {
            int o0 = vo[ev0[e]], o1 = vo[ev1[e]];
            ec_pos[e] = ec_neg[e] = ec_x[e] = NI;
            ec_split[e] = 0;
            if (o0 == 0 && o1 == 0) continue;
            if (o0 >= 0 && o1 >= 0)
                ec_pos[e] = e;
            else if (o0 <= 0 && o1 <= 0)
                ec_neg[e] = e;
            else {
                if (nv + 1 > Caps::MAXV || ne + 2 > Caps::MAXE) {
                    err = 1;
                    return -1;
                }
                ec_split[e] = 1;
                int x = nv++;
                vm[x][0] = em[e][0];
                vm[x][1] = em[e][1];
                vm[x][2] = em[e][2];
                vm[x][3] = (uint8_t)mid;
                vo[x] = 0;
                ec_x[e] = x;
                int a = ne, b = ne + 1;
                ne += 2;
                ec_pos[e] = a;
                ec_neg[e] = b;
                int first = (o0 > 0) ? a : b, second = (o0 > 0) ? b : a;
                ev0[first] = ev0[e];
                ev1[first] = x;
                ev0[second] = x;
                ev1[second] = ev1[e];
                for (int q = 0; q < 3; ++q) em[a][q] = em[b][q] = em[e][q];
                ec_split[a] = ec_split[b] = 0;
            }
        }
        // ---- faces (appended to buffer B)
        const int nF = nf;
        for (int f = 0; f < nF; ++f) {
            const int n = flen[B][f], off = foff[B][f];
            int npos = 0, nneg = 0;
            for (int k = 0; k < n; ++k) {
                int o = vo[fv[B][off + k]];
                npos += (o > 0);
                nneg += (o < 0);
            }
            fc_pos[f] = fc_neg[f] = fc_cut[f] = NI;
            fc_split[f] = 0;
            if (npos == 0 && nneg == 0) continue;
            if (nneg == 0) {
                fc_pos[f] = f;
                continue;
            }
            if (npos == 0) {
                fc_neg[f] = f;
                continue;
            }
            fc_split[f] = 1;
            if (nf + 2 > Caps::MAXF || ne + 1 > Caps::MAXE || nfe + n + 4 > Caps::MAXFE) {
                err = 1;
                return -1;
            }
#define RIN_O(k) ((int)vo[fv[B][off + ((k) % n)]])
            int i = 0;
            while (!(RIN_O(i) <= 0 && RIN_O(i + 1) > 0)) ++i;
            int jl = i + 1;
            while (RIN_O(jl + 1) > 0) ++jl;
            const int ei = fe[B][off + (i % n)], ejl = fe[B][off + (jl % n)];
            int start_tv, end_tv, first_pos, last_pos, first_neg, last_neg;
            if (RIN_O(i) == 0) {
                start_tv = fv[B][off + (i % n)];
                first_pos = ei;
                last_neg = fe[B][off + ((i + n - 1) % n)];
            } else {
                start_tv = ec_x[ei];
                first_pos = ec_pos[ei];
                last_neg = ec_neg[ei];
            }
            if (RIN_O(jl + 1) == 0) {
                end_tv = fv[B][off + ((jl + 1) % n)];
                last_pos = ejl;
                first_neg = fe[B][off + ((jl + 1) % n)];
            } else {
                end_tv = ec_x[ejl];
                last_pos = ec_pos[ejl];
                first_neg = ec_neg[ejl];
            }
            const int ce = ne++;
            ev0[ce] = start_tv;
            ev1[ce] = end_tv;
            {
                int a = (fpos[B][f] == N8) ? fb[B][f] : cmat[fpos[B][f]];
                int b = cmat[fneg[B][f]];
                em[ce][0] = (uint8_t)min(a, b);
                em[ce][1] = (uint8_t)max(a, b);
                em[ce][2] = (uint8_t)mid;
            }
            ec_split[ce] = 0;
            fc_cut[f] = ce;
            const int P = nf, Ng = nf + 1;
            nf += 2;
            fc_pos[f] = P;
            fc_neg[f] = Ng;
            fb[B][P] = fb[B][Ng] = fb[B][f];
            fpos[B][P] = fpos[B][Ng] = fpos[B][f];
            fneg[B][P] = fneg[B][Ng] = fneg[B][f];
            fc_split[P] = fc_split[Ng] = 0;
            fc_cut[P] = fc_cut[Ng] = NI;
            foff[B][P] = nfe;
            fv[B][nfe] = start_tv;
            fe[B][nfe] = first_pos;
            ++nfe;
            for (int k = i + 1; k <= jl; ++k) {
                fv[B][nfe] = fv[B][off + (k % n)];
                fe[B][nfe] = (k == jl) ? last_pos : fe[B][off + (k % n)];
                ++nfe;
            }
            fv[B][nfe] = end_tv;
            fe[B][nfe] = ce;
            ++nfe;
            flen[B][P] = nfe - foff[B][P];
            const int kfirst = (RIN_O(jl + 1) == 0) ? jl + 2 : jl + 1;
            const int klast = (RIN_O(i) == 0) ? i + n - 1 : i + n;
            foff[B][Ng] = nfe;
            fv[B][nfe] = end_tv;
            fe[B][nfe] = first_neg;
            ++nfe;
            for (int k = kfirst; k <= klast; ++k) {
                fv[B][nfe] = fv[B][off + (k % n)];
                fe[B][nfe] = (k == klast) ? last_neg : fe[B][off + (k % n)];
                ++nfe;
            }
            fv[B][nfe] = start_tv;
            fe[B][nfe] = ce;
            ++nfe;
            flen[B][Ng] = nfe - foff[B][Ng];
#undef RIN_O
        }
        // ---- cells: status, cut faces of split cells
        enum { C_NEG = 0, C_POS = 1, C_SPLIT = 2, C_ZERO = 3 };
        const int nC = nc;
        int duplicate_of = -1;
        bool any_pos = false;
        for (int c = 0; c < nC; ++c) {
            bool has_pos = false, has_neg = false;
            for (int f = 0; f < nF; ++f) {
                if (fpos[B][f] != c && fneg[B][f] != c) continue;
                has_pos |= (fc_pos[f] != NI);
                has_neg |= (fc_neg[f] != NI);
            }
            if (!has_pos && !has_neg) {
                cstat[c] = C_ZERO;
                if (duplicate_of < 0) duplicate_of = cmat[c];
                continue;
            }
            if (!has_pos) {
                cstat[c] = C_NEG;
                continue;
            }
            any_pos = true;
            if (!has_neg) {
                cstat[c] = C_POS;
                continue;
            }
            cstat[c] = C_SPLIT;
            I cut_e[Caps::MAXLOOP];
            int n_cut = 0, first_a = -1, first_b = -1;
            auto add_cut_edge = [&](int e, int da, int db, bool inward, bool on_neg_side) {
                for (int k = 0; k < n_cut; ++k)
                    if (cut_e[k] == e) return;
                if (n_cut >= Caps::MAXLOOP) {
                    err = 1;
                    return;
                }
                cut_e[n_cut++] = (I)e;
                if (first_a >= 0) return;
                int oa = inward ? db : da, ob = inward ? da : db;
                if (on_neg_side) {
                    int t = oa;
                    oa = ob;
                    ob = t;
                }
                first_a = oa;
                first_b = ob;
            };
            for (int f = 0; f < nF; ++f) { // the cell's faces in face order
                if (fpos[B][f] != c && fneg[B][f] != c) continue;
                const bool inward = (fpos[B][f] == c);
                if (fc_split[f]) {
                    int ce = fc_cut[f];
                    add_cut_edge(ce, ev0[ce], ev1[ce], inward, true);
                } else if (fc_pos[f] != NI || fc_neg[f] != NI) {
                    const int n = flen[B][f], off = foff[B][f];
                    for (int j = 0; j < n; ++j) {
                        int a = fv[B][off + j], b = fv[B][off + ((j + 1) % n)];
                        if (vo[a] == 0 && vo[b] == 0) add_cut_edge(fe[B][off + j], a, b, inward, fc_neg[f] != NI);
                    }
                }
            }
            if (err) return -1;
            if (nf + 1 > Caps::MAXF || nfe + n_cut > Caps::MAXFE) {
                err = 1;
                return -1;
            }
            const int G = nf++;
            fc_pos[G] = fc_neg[G] = NI;
            fc_split[G] = 0;
            fc_cut[G] = (I)c; // remembers the split cell
            fb[B][G] = N8;
            fpos[B][G] = N8;
            fneg[B][G] = N8;
            foff[B][G] = nfe;
            {
                unsigned long long used = 0;
                int curv = first_a;
                for (int step = 0; step < n_cut; ++step) {
                    int pick = -1;
                    for (int k = 0; k < n_cut; ++k) {
                        if ((used >> k) & 1) continue;
                        int e = cut_e[k];
                        if (step == 0) {
                            if ((ev0[e] == first_a && ev1[e] == first_b) || (ev1[e] == first_a && ev0[e] == first_b)) {
                                pick = k;
                                break;
                            }
                        } else if (ev0[e] == curv || ev1[e] == curv) {
                            pick = k;
                            break;
                        }
                    }
                    if (pick < 0) {
                        err = 2;
                        return -1;
                    }
                    used |= 1ull << pick;
                    int e = cut_e[pick];
                    fv[B][nfe] = (I)curv;
                    fe[B][nfe] = (I)e;
                    ++nfe;
                    curv = (ev0[e] == curv) ? ev1[e] : ev0[e];
                }
                if (curv != first_a) {
                    err = 2;
                    return -1;
                }
            }
            flen[B][G] = n_cut;
        }
        const int n_faces_after_cut = nf;
        // ---- rebuild
        // vertices
        {
            int k = 0;
            for (int v = 0; v < nv; ++v) {
                const bool corner = vm[v][2] < 4;
                if (vo[v] > 0 && !corner) {
                    vmap[v] = NI;
                    continue;
                }
                if (vo[v] > 0) vm[v][3] = (uint8_t)mid;
                vmap[v] = (I)k++;
            }
        }
        // merged simplex-edge pieces
        int n_merged = 0;
        I mg_v0[6], mg_v1[6];
        uint8_t mg_i[6], mg_j[6];
        for (int e = 0; e < ne; ++e) merged_of[e] = NI;
        if (any_pos) {
            for (int i = 0; i < 4; ++i)
                for (int j = i + 1; j < 4; ++j) {
                    int start = -1, end = -1, cnt = 0;
                    for (int e = 0; e < ne; ++e) {
                        if (e < nE && ec_split[e]) continue;
                        if (!(em[e][1] < 4 && em[e][0] == i && em[e][1] == j && edge_is_positive(e, nE))) continue;
                        ++cnt;
                        bool has_pred = false, has_succ = false;
                        for (int g = 0; g < ne; ++g) {
                            if (g < nE && ec_split[g]) continue;
                            if (!(em[g][1] < 4 && em[g][0] == i && em[g][1] == j && edge_is_positive(g, nE))) continue;
                            if (ev1[g] == ev0[e]) has_pred = true;
                            if (ev0[g] == ev1[e]) has_succ = true;
                        }
                        if (!has_pred) start = ev0[e];
                        if (!has_succ) end = ev1[e];
                        merged_of[e] = (I)n_merged;
                    }
                    if (!cnt) continue;
                    mg_v0[n_merged] = (I)start;
                    mg_v1[n_merged] = (I)end;
                    mg_i[n_merged] = (uint8_t)i;
                    mg_j[n_merged] = (uint8_t)j;
                    ++n_merged;
                }
        }
        // edge remap (ec_pos reused): survivors in order, merged edges appended
        int n_surv = 0;
        for (int e = 0; e < ne; ++e) {
            const bool dead = (e < nE && ec_split[e]) || (any_pos && edge_is_positive(e, nE));
            // note: edge_is_positive reads ec_pos[e] for e < nE, so write the remap afterwards
            ec_neg[e] = dead ? NI : (I)n_surv++;
        }
        const int merged_base = n_surv;
        // new edge id of old edge e
        auto map_edge = [&](int e) -> int { return merged_of[e] != NI ? merged_base + merged_of[e] : (int)ec_neg[e]; };
        // cells
        int n_new_cells = 0;
        for (int c = 0; c < nC; ++c) cmap[c] = (cstat[c] == C_NEG || cstat[c] == C_ZERO) ? (uint8_t)n_new_cells++ : N8;
        for (int c = 0; c < nC; ++c) cneg[c] = (cstat[c] == C_SPLIT) ? (uint8_t)n_new_cells++ : N8;
        const int new_cell = any_pos ? n_new_cells++ : N8;
        if (n_new_cells > Caps::MAXC) {
            err = 1;
            return -1;
        }
        uint8_t new_cmat[Caps::MAXC];
        for (int c = 0; c < nC; ++c) {
            if (cmap[c] != N8) new_cmat[cmap[c]] = cmat[c];
            if (cneg[c] != N8) new_cmat[cneg[c]] = cmat[c];
        }
        if (any_pos) new_cmat[new_cell] = (uint8_t)mid;
        auto side_cell = [&](int old_cell) -> int {
            if (old_cell == N8) return N8;
            switch (cstat[old_cell]) {
            case C_NEG:
            case C_ZERO: return cmap[old_cell];
            case C_POS: return new_cell;
            default: return cneg[old_cell];
            }
        };
        // faces -> buffer B2
        int nf2 = 0, nfe2 = 0;
        I bp[4][12];
        int nbp[4] = {0, 0, 0, 0};
        for (int f = 0; f < n_faces_after_cut; ++f) {
            if (f < nF && fc_split[f]) continue;
            const bool is_cut_face = (f >= nF) && fb[B][f] == N8 && fpos[B][f] == N8 && fneg[B][f] == N8;
            bool fpositive = false;
            if (f < nF)
                fpositive = (fc_pos[f] == f);
            else if (!is_cut_face)
                for (int k = 0; k < flen[B][f]; ++k) fpositive |= (vo[fv[B][foff[B][f] + k]] > 0);
            int pc, ncell;
            bool flip = false;
            if (is_cut_face) {
                pc = new_cell;
                ncell = cneg[fc_cut[f]];
            } else {
                if (any_pos && fpositive) {
                    if (fb[B][f] != N8) {
                        if (nbp[fb[B][f]] >= 12) {
                            err = 1;
                            return -1;
                        }
                        bp[fb[B][f]][nbp[fb[B][f]]++] = (I)f;
                    }
                    continue;
                }
                pc = side_cell(fpos[B][f]);
                ncell = side_cell(fneg[B][f]);
                if (fb[B][f] == N8 && pc != N8 && ncell != N8 && new_cmat[pc] < new_cmat[ncell]) {
                    int t = pc;
                    pc = ncell;
                    ncell = t;
                    flip = true;
                }
            }
            const int n = flen[B][f], off = foff[B][f];
            if (nf2 + 1 > Caps::MAXF || nfe2 + n > Caps::MAXFE) {
                err = 1;
                return -1;
            }
            foff[B2][nf2] = nfe2;
            flen[B2][nf2] = n;
            fb[B2][nf2] = fb[B][f];
            fpos[B2][nf2] = (uint8_t)pc;
            fneg[B2][nf2] = (uint8_t)ncell;
            for (int k = 0; k < n; ++k) {
                int sv = flip ? (n - 1 - k) : k;
                int se = flip ? ((2 * n - 2 - k) % n) : k;
                fv[B2][nfe2] = vmap[fv[B][off + sv]];
                fe[B2][nfe2] = (I)map_edge(fe[B][off + se]);
                ++nfe2;
            }
            ++nf2;
        }
        // merged boundary faces
        for (int i = 0; i < 4; ++i) {
            if (!nbp[i]) continue;
            // directed boundary segments of the union (new numbering)
            I sg_from[Caps::MAXLOOP], sg_to[Caps::MAXLOOP], sg_e[Caps::MAXLOOP];
            int ns = 0;
            for (int q = 0; q < nbp[i]; ++q) {
                const int f = bp[i][q], n = flen[B][f], off = foff[B][f];
                for (int k = 0; k < n; ++k) {
                    const int e = fe[B][off + k];
                    const int nee = map_edge(e);
                    if (nee == NI) continue;
                    bool dup = false;
                    for (int s = 0; s < ns; ++s) dup |= (sg_e[s] == nee);
                    if (dup) continue;
                    int a0, a1; // end points of the new edge (old vertex numbering)
                    if (merged_of[e] != NI) {
                        a0 = mg_v0[merged_of[e]];
                        a1 = mg_v1[merged_of[e]];
                    } else {
                        a0 = ev0[e];
                        a1 = ev1[e];
                    }
                    const bool forward = (ev0[e] == fv[B][off + k]);
                    if (ns >= Caps::MAXLOOP) {
                        err = 1;
                        return -1;
                    }
                    sg_from[ns] = vmap[forward ? a0 : a1];
                    sg_to[ns] = vmap[forward ? a1 : a0];
                    sg_e[ns] = (I)nee;
                    ++ns;
                }
            }
            if (nf2 + 1 > Caps::MAXF || nfe2 + ns > Caps::MAXFE) {
                err = 1;
                return -1;
            }
            foff[B2][nf2] = nfe2;
            flen[B2][nf2] = ns;
            fb[B2][nf2] = (uint8_t)i;
            fpos[B2][nf2] = N8;
            fneg[B2][nf2] = (uint8_t)new_cell;
            unsigned long long used = 0;
            int curv = sg_from[0];
            for (int step = 0; step < ns; ++step) {
                int pick = -1;
                for (int k = 0; k < ns; ++k)
                    if (!((used >> k) & 1) && sg_from[k] == curv) {
                        pick = k;
                        break;
                    }
                if (pick < 0) {
                    err = 2;
                    return -1;
                }
                used |= 1ull << pick;
                fv[B2][nfe2] = (I)curv;
                fe[B2][nfe2] = sg_e[pick];
                ++nfe2;
                curv = sg_to[pick];
            }
            if (curv != sg_from[0]) {
                err = 2;
                return -1;
            }
            ++nf2;
        }
        // commit: edges (in place; merged appended), vertices, cells, faces
        {
            for (int e = 0; e < ne; ++e) {
                int d = ec_neg[e];
                if (d == NI) continue;
                ev0[d] = vmap[ev0[e]];
                ev1[d] = vmap[ev1[e]];
                for (int q = 0; q < 3; ++q) em[d][q] = em[e][q];
            }
            if (merged_base + n_merged > Caps::MAXE) {
                err = 1;
                return -1;
            }
            for (int g = 0; g < n_merged; ++g) {
                ev0[merged_base + g] = vmap[mg_v0[g]];
                ev1[merged_base + g] = vmap[mg_v1[g]];
                em[merged_base + g][0] = mg_i[g];
                em[merged_base + g][1] = mg_j[g];
                em[merged_base + g][2] = (uint8_t)mid;
            }
            ne = merged_base + n_merged;
            int k = 0;
            for (int v = 0; v < nv; ++v) {
                if (vmap[v] == NI) continue;
                for (int q = 0; q < 4; ++q) vm[k][q] = vm[v][q];
                ++k;
            }
            nv = k;
            nc = n_new_cells;
            for (int c = 0; c < nc; ++c) cmat[c] = new_cmat[c];
            nf = nf2;
            nfe = nfe2;
            cur = B2;
        }
        return duplicate_of;
    }

    __device__ void insert(const double v[4])
    {
        if (err) return;
        if (nm >= Caps::MAXK + 4) {
            err = 1;
            return;
        }
        int mid = nm++;
        for (int c = 0; c < 4; ++c) mval[mid - 4][c] = v[c];
        int dup = add_material(mid);
        if (err) return;
        if (dup < 0)
            umi[mid] = n_groups++;
        else {
            umi[mid] = umi[dup];
            has_dup = true;
        }
    }

    __device__ bool is_mi_face(int f) const { return fpos[cur][f] != N8; }
    __device__ int pos_label(int f) const { return fpos[cur][f] == N8 ? fb[cur][f] : cmat[fpos[cur][f]]; }
    __device__ int neg_label(int f) const { return cmat[fneg[cur][f]]; }
};

} // namespace rin
