// General per-tet plane arrangement (the ">= 3 functions / degenerate / lookup disabled" path).
//
// Replaces compute_arrangement() of the un-vendored qnzhou/simplicial_arrangement library, called
// by the reference at /root/reference/src/implicit_arrangement.cpp:279,283.  One thread owns one
// tetrahedron; the cell complex lives in fixed-capacity per-thread arrays (local memory, indices
// 8/16 bit).  The algorithm is incremental plane insertion with exact vertex classification
// (predicates.cuh); conventions (ordering, loop orientation, cell sides) are those of DESIGN.md
// "per-tet complex conventions" so that local ids agree with the CPU oracle bit for bit.
#pragma once
#include "predicates.cuh"

#ifndef RIN_TRACK_PEAK
#define RIN_TRACK_PEAK(nv, ne, nf, nc, nfe, ncf) // sizing instrumentation hook (scripts/size_caps.cpp)
#endif

namespace rin {

// Big tier: per-thread local memory, 16-bit indices.
struct IACaps
{
    using idx = uint16_t;
    static constexpr int MAXK = 64;   // input planes
    static constexpr int MAXV = 256;
    static constexpr int MAXE = 640;
    static constexpr int MAXF = 448;
    static constexpr int MAXC = 160;
    static constexpr int MAXFE = 2048; // face loop pool
    static constexpr int MAXCF = 1024; // cell face pool
    static constexpr int MAXLOOP = 48;
};
// Small tier (<= 4 functions, the common non-table case): 8-bit indices, ~3.9 KB, lives in
// shared memory.  Capacities cover the transient state inside one insertion step.
struct IACapsSmall
{
    using idx = uint8_t;
    static constexpr int MAXK = 4;
    static constexpr int MAXV = 48;
    static constexpr int MAXE = 128;
    static constexpr int MAXF = 96;
    static constexpr int MAXC = 32;
    static constexpr int MAXFE = 384;
    static constexpr int MAXCF = 160;
    static constexpr int MAXLOOP = 12;
};

// Mid tier (5..20 functions, or a small-tier overflow): 8-bit indices, ~9 KB, shared memory, one tet per
// warp.  Sized from the peak transient counts of the 32-function stress configuration (scripts/size_caps.cpp:
// ne <= 239, nf <= 195, nfe <= 779 at k = 20); anything larger falls through to the big tier.
struct IACapsMid
{
    using idx = uint8_t;
    static constexpr int MAXK = 20;
    static constexpr int MAXV = 128;
    static constexpr int MAXE = 254;
    static constexpr int MAXF = 224;
    static constexpr int MAXC = 64;
    static constexpr int MAXFE = 1024;
    static constexpr int MAXCF = 384;
    static constexpr int MAXLOOP = 24;
};

constexpr uint8_t N8 = 0xff;

template <class Caps>
struct IAComplex
{
    using I = typename Caps::idx;
    static constexpr I NI = (I)~(I)0; // "none"
    // geometry: input plane j (plane id 4+j) -> values at the 4 simplex corners
    double plv[Caps::MAXK][4];
    int np;
    // combinatorics
    int nv, ne, nf, nc, nfe, ncf;
    uint8_t vp[Caps::MAXV][3];
    int8_t vo[Caps::MAXV];
    I ev0[Caps::MAXE], ev1[Caps::MAXE];
    uint8_t ep0[Caps::MAXE], ep1[Caps::MAXE];
    I ec_pos[Caps::MAXE], ec_neg[Caps::MAXE], ec_x[Caps::MAXE]; // also reused as edge remap
    uint8_t ec_split[Caps::MAXE];
    uint16_t foff[Caps::MAXF];
    uint8_t flen[Caps::MAXF], fplane[Caps::MAXF], fpos[Caps::MAXF], fneg[Caps::MAXF];
    I fc_pos[Caps::MAXF], fc_neg[Caps::MAXF], fc_cut[Caps::MAXF]; // fc_pos reused as face remap
    uint8_t fc_split[Caps::MAXF];
    I fv[Caps::MAXFE], fe[Caps::MAXFE];
    uint16_t coff[Caps::MAXC];
    uint8_t clen[Caps::MAXC], c_split[Caps::MAXC], cmap[Caps::MAXC];
    I cf[Caps::MAXCF];
    // unique-plane bookkeeping
    uint8_t upi[Caps::MAXK + 4]; // plane -> group
    int n_groups;
    bool has_coplanar;
    int err; // 0 ok, 1 capacity, 2 degenerate input
    unsigned n_exact;

    __device__ void init()
    {
        np = 4;
        err = 0;
        n_exact = 0;
        nv = 4;
        const uint8_t vps[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) vp[i][k] = vps[i][k];
        // edges (a,b), a<b, lexicographic: 01 02 03 12 13 23
        const uint8_t ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};
        const uint8_t epa[6] = {2, 1, 1, 0, 0, 0}, epb[6] = {3, 3, 2, 3, 2, 1};
        ne = 6;
        for (int e = 0; e < 6; ++e) {
            ev0[e] = ea[e];
            ev1[e] = eb[e];
            ep0[e] = epa[e];
            ep1[e] = epb[e];
        }
        // faces: loops CCW seen from inside; eid[a][b]
        const uint8_t eid[4][4] = {{0, 0, 1, 2}, {0, 0, 3, 4}, {1, 3, 0, 5}, {2, 4, 5, 0}};
        const uint8_t loops[4][3] = {{1, 3, 2}, {0, 2, 3}, {0, 3, 1}, {0, 1, 2}};
        nf = 4;
        nfe = 0;
        for (int f = 0; f < 4; ++f) {
            foff[f] = nfe;
            flen[f] = 3;
            fplane[f] = f;
            fpos[f] = 0;
            fneg[f] = N8;
            for (int k = 0; k < 3; ++k) {
                fv[nfe] = loops[f][k];
                fe[nfe] = eid[loops[f][k]][loops[f][(k + 1) % 3]];
                ++nfe;
            }
        }
        nc = 1;
        coff[0] = 0;
        clen[0] = 4;
        ncf = 4;
        for (int f = 0; f < 4; ++f) cf[f] = f;
        for (int p = 0; p < 4; ++p) upi[p] = p;
        n_groups = 4;
        has_coplanar = false;
    }

    // exact sign of plane q at the point where the three planes of vertex v meet:
    // q(X) = det[impl; q] / det[impl; 1] restricted to the corners not fixed by boundary planes
    // nex / bad: exact-fallback counter and degeneracy flag of the caller (per lane in the warp version)
    __device__ int orient_vertex(int v, const double* q, unsigned& nex, int& bad)
    {
        const double* impl[3];
        int ni = 0;
        unsigned fixed = 0;
        for (int k = 0; k < 3; ++k) {
            int p = vp[v][k];
            if (p < 4)
                fixed |= 1u << p;
            else
                impl[ni++] = plv[p - 4];
        }
        int idx[4], n = 0;
        for (int c = 0; c < 4; ++c)
            if (!((fixed >> c) & 1)) idx[n++] = c;
        if (n == 1) {
            double x = q[idx[0]];
            return x > 0 ? 1 : (x < 0 ? -1 : 0);
        }
        int sq, sd;
        if (n == 2) {
            const double a = impl[0][idx[0]], b = impl[0][idx[1]];
            sq = det2_sign(a, b, q[idx[0]], q[idx[1]], &nex);
            if (sq == 0) return 0;
            sd = (a > b) - (a < b); // det[[a,b],[1,1]]
        } else if (n == 3) {
            double m[9];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                m[c] = impl[0][idx[c]];
                m[3 + c] = impl[1][idx[c]];
                m[6 + c] = q[idx[c]];
            }
            sq = det3_sign(m, &nex);
            if (sq == 0) return 0;
            m[6] = m[7] = m[8] = 1.0;
            sd = det3_sign(m, &nex);
        } else {
            double m[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                m[c] = impl[0][c];
                m[4 + c] = impl[1][c];
                m[8 + c] = impl[2][c];
                m[12 + c] = q[c];
            }
            sq = det4_sign(m, &nex);
            if (sq == 0) return 0;
            m[12] = m[13] = m[14] = m[15] = 1.0;
            sd = det4_sign(m, &nex);
        }
        if (sd == 0) bad = 2;
        return sq * sd;
    }

    // position k of edge e in the loop of face f (k with fe[foff+k]==e)
    __device__ int edge_pos_in_face(int f, int e) const
    {
        for (int k = 0; k < flen[f]; ++k)
            if (fe[foff[f] + k] == e) return k;
        return -1;
    }

    // insert plane `pid` (values already stored in pl[pid]); returns coplanar plane id or -1
    __device__ int add_plane(int pid)
    {
        const double* q = plv[pid - 4];
        // ---- 1. vertices
        bool any = false;
        for (int v = 0; v < nv; ++v) {
            vo[v] = (int8_t)orient_vertex(v, q, n_exact, err);
            any |= (vo[v] != 0);
        }
        if (!any) {
            err = 2;
            return -1;
        }
        if (err) return -1;
        // ---- 2. edges
        const int nE = ne;
        for (int e = 0; e < nE; ++e) {
            int o0 = vo[ev0[e]], o1 = vo[ev1[e]];
            ec_pos[e] = ec_neg[e] = ec_x[e] = NI;
            ec_split[e] = 0;
            if (o0 == 0 && o1 == 0) continue;
            if (o0 == 0)
                ec_x[e] = ev0[e];
            else if (o1 == 0)
                ec_x[e] = ev1[e];
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
                vp[x][0] = ep0[e];
                vp[x][1] = ep1[e];
                vp[x][2] = (uint8_t)pid;
                vo[x] = 0;
                ec_x[e] = x;
                // sub-edges keep the direction v0 -> x -> v1; positive one is stored first
                int a = ne, b = ne + 1;
                ne += 2;
                int pe = a, ng = b;
                ec_pos[e] = pe;
                ec_neg[e] = ng;
                int first = (o0 > 0) ? pe : ng, second = (o0 > 0) ? ng : pe;
                ev0[first] = ev0[e];
                ev1[first] = x;
                ev0[second] = x;
                ev1[second] = ev1[e];
                ep0[a] = ep0[b] = ep0[e];
                ep1[a] = ep1[b] = ep1[e];
                ec_split[a] = ec_split[b] = 0;
            }
        }
        // ---- 3. faces
        const int nF = nf;
        int coplanar = -1;
        for (int f = 0; f < nF; ++f) {
            const int n = flen[f], off = foff[f];
            int npos = 0, nneg = 0;
            for (int k = 0; k < n; ++k) {
                int o = vo[fv[off + k]];
                npos += (o > 0);
                nneg += (o < 0);
            }
            fc_pos[f] = fc_neg[f] = fc_cut[f] = NI;
            fc_split[f] = 0;
            if (npos == 0 && nneg == 0) {
                if (coplanar < 0) coplanar = fplane[f];
                continue;
            }
            if (nneg == 0) {
                fc_pos[f] = f;
                continue;
            }
            if (npos == 0) {
                fc_neg[f] = f;
                continue;
            }
            fc_split[f] = 1;
            if (nf + 2 > Caps::MAXF || ne + 1 > Caps::MAXE || nfe + 2 * n + 4 > Caps::MAXFE) {
                err = 1;
                return -1;
            }
#define RIN_O(k) ((int)vo[fv[off + ((k) % n)]])
            int i = 0;
            while (!(RIN_O(i) <= 0 && RIN_O(i + 1) > 0)) ++i;
            int jl = i + 1;
            while (RIN_O(jl + 1) > 0) ++jl;
            const int ei = fe[off + (i % n)], ejl = fe[off + (jl % n)];
            int start_tv, end_tv, first_pos, last_pos, first_neg, last_neg;
            if (RIN_O(i) == 0) {
                start_tv = fv[off + (i % n)];
                first_pos = ei;
                last_neg = fe[off + ((i + n - 1) % n)];
            } else {
                start_tv = ec_x[ei];
                first_pos = ec_pos[ei];
                last_neg = ec_neg[ei];
            }
            if (RIN_O(jl + 1) == 0) {
                end_tv = fv[off + ((jl + 1) % n)];
                last_pos = ejl;
                first_neg = fe[off + ((jl + 1) % n)];
            } else {
                end_tv = ec_x[ejl];
                last_pos = ec_pos[ejl];
                first_neg = ec_neg[ejl];
            }
            const int ce = ne++;
            ev0[ce] = start_tv;
            ev1[ce] = end_tv;
            ep0[ce] = fplane[f];
            ep1[ce] = (uint8_t)pid;
            ec_split[ce] = 0;
            fc_cut[f] = ce;
            // positive loop
            const int P = nf, Ng = nf + 1;
            nf += 2;
            fc_pos[f] = P;
            fc_neg[f] = Ng;
            fplane[P] = fplane[Ng] = fplane[f];
            fpos[P] = fpos[Ng] = fpos[f];
            fneg[P] = fneg[Ng] = fneg[f];
            fc_split[P] = fc_split[Ng] = 0;
            foff[P] = nfe;
            fv[nfe] = start_tv;
            fe[nfe] = first_pos;
            ++nfe;
            for (int k = i + 1; k <= jl; ++k) {
                fv[nfe] = fv[off + (k % n)];
                fe[nfe] = (k == jl) ? last_pos : fe[off + (k % n)];
                ++nfe;
            }
            fv[nfe] = end_tv;
            fe[nfe] = ce;
            ++nfe;
            flen[P] = nfe - foff[P];
            // negative loop
            const int kfirst = (RIN_O(jl + 1) == 0) ? jl + 2 : jl + 1;
            const int klast = (RIN_O(i) == 0) ? i + n - 1 : i + n;
            foff[Ng] = nfe;
            fv[nfe] = end_tv;
            fe[nfe] = first_neg;
            ++nfe;
            for (int k = kfirst; k <= klast; ++k) {
                fv[nfe] = fv[off + (k % n)];
                fe[nfe] = (k == klast) ? last_neg : fe[off + (k % n)];
                ++nfe;
            }
            fv[nfe] = start_tv;
            fe[nfe] = ce;
            ++nfe;
            flen[Ng] = nfe - foff[Ng];
#undef RIN_O
        }
        // ---- 4. cells
        const int nC = nc;
        for (int c = 0; c < nC; ++c) c_split[c] = 0;
        for (int c = 0; c < nC; ++c) {
            bool has_pos = false, has_neg = false;
            for (int k = 0; k < clen[c]; ++k) {
                int f = cf[coff[c] + k];
                has_pos |= (fc_pos[f] != NI);
                has_neg |= (fc_neg[f] != NI);
            }
            if (!(has_pos && has_neg)) continue;
            c_split[c] = 1;
            if (nc + 2 > Caps::MAXC || nf + 1 > Caps::MAXF || ncf + 2 * clen[c] + 2 > Caps::MAXCF) {
                err = 1;
                return -1;
            }
            // collect the boundary edges of the cut polygon
            I cut_e[Caps::MAXLOOP];
            int n_cut = 0;
            int first_a = -1, first_b = -1;
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
            const int cp = nc, cn = nc + 1;
            nc += 2;
            c_split[cp] = c_split[cn] = 0;
            // positive sub-cell's faces, then the negative one's (lists are contiguous in the pool)
            coff[cp] = ncf;
            for (int k = 0; k < clen[c]; ++k) {
                int f = cf[coff[c] + k];
                if (fc_pos[f] != NI) cf[ncf++] = fc_pos[f];
            }
            const int gpos_slot = ncf++;
            clen[cp] = ncf - coff[cp];
            coff[cn] = ncf;
            for (int k = 0; k < clen[c]; ++k) {
                int f = cf[coff[c] + k];
                if (fc_neg[f] != NI) cf[ncf++] = fc_neg[f];
            }
            const int gneg_slot = ncf++;
            clen[cn] = ncf - coff[cn];
            for (int k = 0; k < clen[c]; ++k) {
                int f = cf[coff[c] + k];
                const bool inward = (fpos[f] == c);
                if (fc_split[f]) {
                    int ce = fc_cut[f];
                    add_cut_edge(ce, ev0[ce], ev1[ce], inward, true);
                } else if (fc_pos[f] != NI || fc_neg[f] != NI) {
                    const int n = flen[f], off = foff[f];
                    for (int j = 0; j < n; ++j) {
                        int a = fv[off + j], b = fv[off + ((j + 1) % n)];
                        if (vo[a] == 0 && vo[b] == 0)
                            add_cut_edge(fe[off + j], a, b, inward, fc_neg[f] != NI);
                    }
                }
            }
            if (err) return -1;
            // new face G: chain the cut edges into a loop starting first_a -> first_b
            const int G = nf++;
            fc_pos[G] = fc_neg[G] = fc_cut[G] = NI;
            fc_split[G] = 0;
            fplane[G] = (uint8_t)pid;
            fpos[G] = (uint8_t)cp;
            fneg[G] = (uint8_t)cn;
            foff[G] = nfe;
            if (nfe + n_cut > Caps::MAXFE) {
                err = 1;
                return -1;
            }
            {
                unsigned long long used = 0;
                int cur = first_a;
                for (int step = 0; step < n_cut; ++step) {
                    int pick = -1;
                    for (int k = 0; k < n_cut; ++k) {
                        if ((used >> k) & 1) continue;
                        int e = cut_e[k];
                        if (step == 0) {
                            if ((ev0[e] == first_a && ev1[e] == first_b) ||
                                (ev1[e] == first_a && ev0[e] == first_b)) {
                                pick = k;
                                break;
                            }
                        } else if (ev0[e] == cur || ev1[e] == cur) {
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
                    fv[nfe] = (I)cur;
                    fe[nfe] = (I)e;
                    ++nfe;
                    cur = (ev0[e] == cur) ? ev1[e] : ev0[e];
                }
                if (cur != first_a) {
                    err = 2;
                    return -1;
                }
            }
            flen[G] = n_cut;
            cf[gpos_slot] = G;
            cf[gneg_slot] = G;
            for (int k = 0; k < clen[cp] - 1; ++k) {
                int f = cf[coff[cp] + k];
                if (fpos[f] == c) fpos[f] = (uint8_t)cp;
                if (fneg[f] == c) fneg[f] = (uint8_t)cp;
            }
            for (int k = 0; k < clen[cn] - 1; ++k) {
                int f = cf[coff[cn] + k];
                if (fpos[f] == c) fpos[f] = (uint8_t)cn;
                if (fneg[f] == c) fneg[f] = (uint8_t)cn;
            }
        }
        // ---- 5. consolidate in place (survivors keep their order)
        RIN_TRACK_PEAK(nv, ne, nf, nc, nfe, ncf);
        {
            int k = 0;
            for (int e = 0; e < ne; ++e) {
                bool dead = (e < nE) && ec_split[e];
                ec_pos[e] = dead ? NI : (I)k; // reuse as edge remap
                if (!dead) {
                    ev0[k] = ev0[e];
                    ev1[k] = ev1[e];
                    ep0[k] = ep0[e];
                    ep1[k] = ep1[e];
                    ++k;
                }
            }
            ne = k;
            int kc = 0;
            for (int c = 0; c < nc; ++c) cmap[c] = ((c < nC) && c_split[c]) ? N8 : (uint8_t)kc++;
            int kf = 0, pool = 0;
            for (int f = 0; f < nf; ++f) {
                bool dead = (f < nF) && fc_split[f];
                if (dead) {
                    fc_pos[f] = NI;
                    continue;
                }
                int off = foff[f], n = flen[f];
                foff[kf] = pool;
                flen[kf] = n;
                fplane[kf] = fplane[f];
                fpos[kf] = (fpos[f] == N8) ? N8 : cmap[fpos[f]];
                fneg[kf] = (fneg[f] == N8) ? N8 : cmap[fneg[f]];
                for (int j = 0; j < n; ++j) {
                    fv[pool] = fv[off + j];
                    fe[pool] = ec_pos[fe[off + j]];
                    ++pool;
                }
                fc_pos[f] = kf++; // reuse as face remap (kf <= f, so the slot f is already consumed)
            }
            nf = kf;
            nfe = pool;
            int cpool = 0;
            for (int c = 0; c < nc; ++c) {
                if (cmap[c] == N8) continue;
                int off = coff[c], n = clen[c], d = cmap[c];
                coff[d] = cpool;
                clen[d] = n;
                for (int j = 0; j < n; ++j) cf[cpool++] = fc_pos[cf[off + j]];
            }
            nc = kc;
            ncf = cpool;
        }
        return coplanar;
    }

    // add the next input plane (values at the 4 corners); tracks coincident planes
    __device__ void insert(const double v[4])
    {
        if (err) return;
        if (np >= Caps::MAXK + 4) {
            err = 1;
            return;
        }
        int pid = np++;
        for (int c = 0; c < 4; ++c) plv[pid - 4][c] = v[c];
        int cop = add_plane(pid);
        if (err) return;
        if (cop < 0)
            upi[pid] = n_groups++;
        else {
            upi[pid] = upi[cop];
            has_coplanar = true;
        }
    }

    // first plane (lowest id) of the group of plane p
    __device__ int group_first(int p) const
    {
        for (int r = 0; r < np; ++r)
            if (upi[r] == upi[p]) return r;
        return p;
    }

    __device__ bool same_orientation(int p) const
    {
        int r = group_first(p);
        if (r == p) return true;
        // p = c * r as linear functions; r < 4 is the unit plane e_r
        if (r < 4) return plv[p - 4][r] > 0;
        for (int k = 0; k < 4; ++k)
            if (plv[r - 4][k] != 0) return (plv[r - 4][k] > 0) == (plv[p - 4][k] > 0);
        return true;
    }

    // face f lies on an isosurface: its supporting plane or a coincident one is an input plane
    // (/root/reference/src/extract_mesh.cpp:65-91)
    __device__ bool is_iso_face(int f) const
    {
        int sp = fplane[f];
        if (sp > 3) return true;
        if (!has_coplanar) return false;
        for (int p = 4; p < np; ++p)
            if (upi[p] == upi[sp]) return true;
        return false;
    }
};

} // namespace rin
