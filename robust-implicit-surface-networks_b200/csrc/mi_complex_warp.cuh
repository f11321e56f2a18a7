// Warp-cooperative material insertion: one tetrahedron per warp, the complex in shared memory.
//
// Same algorithm and the same numbering as MIComplex::add_material (mi_complex.cuh), which replaces
// compute_material_interface() of the un-vendored simplicial_arrangement library
// (/root/reference/src/material_interface.cpp:323,327).  Lanes own the entities of a 32-wide chunk;
// ids of new / surviving entities are "running count + rank among the lanes before me", which is the
// id the serial loops assign.  The few inherently small loops (6 simplex edges, 4 simplex faces) run
// on 6 / 4 lanes.  All 32 lanes call every function with identical arguments.
// tests/simt/ runs this file on the CPU (32 threads per warp) against the serial version.
#pragma once
#include "ia_complex_warp.cuh"
#include "mi_complex.cuh"

namespace rin {

template <class Caps>
struct MIWarpScratch
{
    using I = typename Caps::idx;
    uint8_t new_cmat[Caps::MAXC];
    I mg_v0[6], mg_v1[6];
    uint8_t mg_i[6], mg_j[6];
    I bp[4][12];
};

template <class Caps>
__device__ __forceinline__ int mi_warp_fail(MIComplex<Caps>& cx, int code, int lane)
{
    if (lane == 0) cx.err = code;
    __syncwarp();
    return -1;
}

// inserts material `mid` (values in mval[mid-4]); returns a material it duplicates, or -1
template <class Caps>
__device__ int warp_add_material(MIComplex<Caps>& cx, MIWarpScratch<Caps>& sc, int mid, int lane)
{
    using I = typename Caps::idx;
    constexpr I NI = MIComplex<Caps>::NI;
    enum { C_NEG = 0, C_POS = 1, C_SPLIT = 2, C_ZERO = 3 };
    const unsigned lt = (1u << lane) - 1u;
    const double* M = cx.mval[mid - 4];
    const int B = cx.cur, B2 = cx.cur ^ 1;
    int nv = cx.nv, ne = cx.ne, nf = cx.nf, nfe = cx.nfe;
    const int nC = cx.nc;

    // ---- vertices
    {
        unsigned nex = 0;
        int bad = 0;
        for (int v = lane; v < nv; v += 32) cx.vo[v] = (int8_t)cx.orient_vertex(v, M, nex, bad);
        const bool degenerate = __ballot_sync(WFULL, bad != 0) != 0u;
        if (__ballot_sync(WFULL, nex != 0u)) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) nex += __shfl_xor_sync(WFULL, nex, d);
            if (lane == 0) cx.n_exact += nex;
        }
        if (degenerate) return mi_warp_fail(cx, 2, lane);
        __syncwarp();
    }

    // ---- edges
    const int nE = ne;
    for (int base = 0; base < nE; base += 32) {
        const int e = base + lane;
        bool split = false;
        int o0 = 0;
        if (e < nE) {
            o0 = cx.vo[cx.ev0[e]];
            const int o1 = cx.vo[cx.ev1[e]];
            I p = NI, n = NI;
            if (!(o0 == 0 && o1 == 0)) {
                if (o0 >= 0 && o1 >= 0)
                    p = (I)e;
                else if (o0 <= 0 && o1 <= 0)
                    n = (I)e;
                else
                    split = true;
            }
            cx.ec_pos[e] = p;
            cx.ec_neg[e] = n;
            cx.ec_x[e] = NI;
            cx.ec_split[e] = split ? 1 : 0;
        }
        const unsigned ms = __ballot_sync(WFULL, split);
        if (!ms) continue;
        const int cnt = __popc(ms);
        if (nv + cnt > Caps::MAXV || ne + 2 * cnt > Caps::MAXE) return mi_warp_fail(cx, 1, lane);
        if (split) {
            const int r = __popc(ms & lt);
            const int x = nv + r, a = ne + 2 * r, b = a + 1;
            cx.vm[x][0] = cx.em[e][0];
            cx.vm[x][1] = cx.em[e][1];
            cx.vm[x][2] = cx.em[e][2];
            cx.vm[x][3] = (uint8_t)mid;
            cx.vo[x] = 0;
            cx.ec_x[e] = (I)x;
            cx.ec_pos[e] = (I)a;
            cx.ec_neg[e] = (I)b;
            const int first = (o0 > 0) ? a : b, second = (o0 > 0) ? b : a;
            cx.ev0[first] = cx.ev0[e];
            cx.ev1[first] = (I)x;
            cx.ev0[second] = (I)x;
            cx.ev1[second] = cx.ev1[e];
            for (int q = 0; q < 3; ++q) cx.em[a][q] = cx.em[b][q] = cx.em[e][q];
            cx.ec_split[a] = cx.ec_split[b] = 0;
        }
        nv += cnt;
        ne += 2 * cnt;
    }
    __syncwarp();

    // ---- faces (appended to buffer B)
    const int nF = nf;
    for (int base = 0; base < nF; base += 32) {
        const int f = base + lane;
        int kind = 0; // 1 on the boundary of M's region, 2 positive, 3 negative, 4 split
        int n = 0, off = 0;
        if (f < nF) {
            n = cx.flen[B][f];
            off = cx.foff[B][f];
            int npos = 0, nneg = 0;
            for (int k = 0; k < n; ++k) {
                const int o = cx.vo[cx.fv[B][off + k]];
                npos += (o > 0);
                nneg += (o < 0);
            }
            kind = (npos == 0 && nneg == 0) ? 1 : (nneg == 0 ? 2 : (npos == 0 ? 3 : 4));
            cx.fc_pos[f] = (kind == 2) ? (I)f : NI;
            cx.fc_neg[f] = (kind == 3) ? (I)f : NI;
            cx.fc_cut[f] = NI;
            cx.fc_split[f] = (kind == 4) ? 1 : 0;
        }
        const unsigned ms = __ballot_sync(WFULL, kind == 4);
        if (!ms) continue;
#define RIN_O(k) ((int)cx.vo[cx.fv[B][off + ((k) % n)]])
        int i = 0, jl = 0, kfirst = 0, klast = 0, need = 0;
        if (kind == 4) {
            while (!(RIN_O(i) <= 0 && RIN_O(i + 1) > 0)) ++i;
            jl = i + 1;
            while (RIN_O(jl + 1) > 0) ++jl;
            kfirst = (RIN_O(jl + 1) == 0) ? jl + 2 : jl + 1;
            klast = (RIN_O(i) == 0) ? i + n - 1 : i + n;
            need = (jl - i + 2) + (klast - kfirst + 3);
        }
        int tot;
        const int poff = warp_excl_scan(need, lane, tot);
        const int cnt = __popc(ms);
        if (nf + 2 * cnt > Caps::MAXF || ne + cnt > Caps::MAXE || nfe + tot > Caps::MAXFE)
            return mi_warp_fail(cx, 1, lane);
        if (kind == 4) {
            const int r = __popc(ms & lt);
            const int ei = cx.fe[B][off + (i % n)], ejl = cx.fe[B][off + (jl % n)];
            int start_tv, end_tv, first_pos, last_pos, first_neg, last_neg;
            if (RIN_O(i) == 0) {
                start_tv = cx.fv[B][off + (i % n)];
                first_pos = ei;
                last_neg = cx.fe[B][off + ((i + n - 1) % n)];
            } else {
                start_tv = cx.ec_x[ei];
                first_pos = cx.ec_pos[ei];
                last_neg = cx.ec_neg[ei];
            }
            if (RIN_O(jl + 1) == 0) {
                end_tv = cx.fv[B][off + ((jl + 1) % n)];
                last_pos = ejl;
                first_neg = cx.fe[B][off + ((jl + 1) % n)];
            } else {
                end_tv = cx.ec_x[ejl];
                last_pos = cx.ec_pos[ejl];
                first_neg = cx.ec_neg[ejl];
            }
            const int ce = ne + r;
            cx.ev0[ce] = (I)start_tv;
            cx.ev1[ce] = (I)end_tv;
            {
                const int a = (cx.fpos[B][f] == N8) ? (int)cx.fb[B][f] : (int)cx.cmat[cx.fpos[B][f]];
                const int b = cx.cmat[cx.fneg[B][f]];
                cx.em[ce][0] = (uint8_t)min(a, b);
                cx.em[ce][1] = (uint8_t)max(a, b);
                cx.em[ce][2] = (uint8_t)mid;
            }
            cx.ec_split[ce] = 0;
            cx.fc_cut[f] = (I)ce;
            const int P = nf + 2 * r, Ng = P + 1;
            cx.fc_pos[f] = (I)P;
            cx.fc_neg[f] = (I)Ng;
            cx.fb[B][P] = cx.fb[B][Ng] = cx.fb[B][f];
            cx.fpos[B][P] = cx.fpos[B][Ng] = cx.fpos[B][f];
            cx.fneg[B][P] = cx.fneg[B][Ng] = cx.fneg[B][f];
            cx.fc_split[P] = cx.fc_split[Ng] = 0;
            cx.fc_cut[P] = cx.fc_cut[Ng] = NI;
            int w = nfe + poff;
            cx.foff[B][P] = (uint16_t)w;
            cx.fv[B][w] = (I)start_tv;
            cx.fe[B][w] = (I)first_pos;
            ++w;
            for (int k = i + 1; k <= jl; ++k) {
                cx.fv[B][w] = cx.fv[B][off + (k % n)];
                cx.fe[B][w] = (k == jl) ? (I)last_pos : cx.fe[B][off + (k % n)];
                ++w;
            }
            cx.fv[B][w] = (I)end_tv;
            cx.fe[B][w] = (I)ce;
            ++w;
            cx.flen[B][P] = (uint8_t)(w - cx.foff[B][P]);
            cx.foff[B][Ng] = (uint16_t)w;
            cx.fv[B][w] = (I)end_tv;
            cx.fe[B][w] = (I)first_neg;
            ++w;
            for (int k = kfirst; k <= klast; ++k) {
                cx.fv[B][w] = cx.fv[B][off + (k % n)];
                cx.fe[B][w] = (k == klast) ? (I)last_neg : cx.fe[B][off + (k % n)];
                ++w;
            }
            cx.fv[B][w] = (I)start_tv;
            cx.fe[B][w] = (I)ce;
            ++w;
            cx.flen[B][Ng] = (uint8_t)(w - cx.foff[B][Ng]);
        }
#undef RIN_O
        ne += cnt;
        nf += 2 * cnt;
        nfe += tot;
    }
    __syncwarp();

    // ---- cells: status, cut faces of split cells (a lane per cell; the cell's faces in face order)
    int duplicate_of = -1;
    bool any_pos = false;
    for (int base = 0; base < nC; base += 32) {
        const int c = base + lane;
        int stat = -1, n_cut = 0, first_a = -1, first_b = -1, lerr = 0;
        I cut_e[Caps::MAXLOOP];
        if (c < nC) {
            bool has_pos = false, has_neg = false;
            for (int f = 0; f < nF; ++f) {
                if (cx.fpos[B][f] != c && cx.fneg[B][f] != c) continue;
                has_pos |= (cx.fc_pos[f] != NI);
                has_neg |= (cx.fc_neg[f] != NI);
            }
            stat = (!has_pos && !has_neg) ? C_ZERO : (!has_pos ? C_NEG : (!has_neg ? C_POS : C_SPLIT));
            cx.cstat[c] = (uint8_t)stat;
            if (stat == C_SPLIT) {
                auto add_cut_edge = [&](int e, int da, int db, bool inward, bool on_neg_side) {
                    for (int k = 0; k < n_cut; ++k)
                        if (cut_e[k] == e) return;
                    if (n_cut >= Caps::MAXLOOP) {
                        lerr = 1;
                        return;
                    }
                    cut_e[n_cut++] = (I)e;
                    if (first_a >= 0) return;
                    int oa = inward ? db : da, ob = inward ? da : db;
                    if (on_neg_side) {
                        const int t = oa;
                        oa = ob;
                        ob = t;
                    }
                    first_a = oa;
                    first_b = ob;
                };
                for (int f = 0; f < nF; ++f) {
                    if (cx.fpos[B][f] != c && cx.fneg[B][f] != c) continue;
                    const bool inward = (cx.fpos[B][f] == c);
                    if (cx.fc_split[f]) {
                        const int ce = cx.fc_cut[f];
                        add_cut_edge(ce, cx.ev0[ce], cx.ev1[ce], inward, true);
                    } else if (cx.fc_pos[f] != NI || cx.fc_neg[f] != NI) {
                        const int n = cx.flen[B][f], off = cx.foff[B][f];
                        for (int j = 0; j < n; ++j) {
                            const int a = cx.fv[B][off + j], b = cx.fv[B][off + ((j + 1) % n)];
                            if (cx.vo[a] == 0 && cx.vo[b] == 0)
                                add_cut_edge(cx.fe[B][off + j], a, b, inward, cx.fc_neg[f] != NI);
                        }
                    }
                }
            }
        }
        const unsigned mz = __ballot_sync(WFULL, stat == C_ZERO);
        if (mz && duplicate_of < 0) {
            const int m = (stat == C_ZERO) ? (int)cx.cmat[c] : 0;
            duplicate_of = __shfl_sync(WFULL, m, __ffs(mz) - 1);
        }
        any_pos |= __ballot_sync(WFULL, stat == C_POS || stat == C_SPLIT) != 0u;
        const unsigned ms = __ballot_sync(WFULL, stat == C_SPLIT);
        if (!ms) continue;
        if (__ballot_sync(WFULL, lerr != 0)) return mi_warp_fail(cx, 1, lane);
        int tot;
        const int feo = warp_excl_scan(stat == C_SPLIT ? n_cut : 0, lane, tot);
        const int cnt = __popc(ms);
        if (nf + cnt > Caps::MAXF || nfe + tot > Caps::MAXFE) return mi_warp_fail(cx, 1, lane);
        int chain_bad = 0;
        if (stat == C_SPLIT) {
            const int G = nf + __popc(ms & lt);
            cx.fc_pos[G] = cx.fc_neg[G] = NI;
            cx.fc_split[G] = 0;
            cx.fc_cut[G] = (I)c; // remembers the split cell
            cx.fb[B][G] = N8;
            cx.fpos[B][G] = N8;
            cx.fneg[B][G] = N8;
            cx.foff[B][G] = (uint16_t)(nfe + feo);
            cx.flen[B][G] = (uint8_t)n_cut;
            int wf = nfe + feo;
            unsigned long long used = 0;
            int curv = first_a;
            for (int step = 0; step < n_cut; ++step) {
                int pick = -1;
                for (int k = 0; k < n_cut; ++k) {
                    if ((used >> k) & 1) continue;
                    const int e = cut_e[k];
                    if (step == 0) {
                        if ((cx.ev0[e] == first_a && cx.ev1[e] == first_b) ||
                            (cx.ev1[e] == first_a && cx.ev0[e] == first_b)) {
                            pick = k;
                            break;
                        }
                    } else if (cx.ev0[e] == curv || cx.ev1[e] == curv) {
                        pick = k;
                        break;
                    }
                }
                if (pick < 0) {
                    chain_bad = 1;
                    break;
                }
                used |= 1ull << pick;
                const int e = cut_e[pick];
                cx.fv[B][wf] = (I)curv;
                cx.fe[B][wf] = (I)e;
                ++wf;
                curv = (cx.ev0[e] == curv) ? cx.ev1[e] : cx.ev0[e];
            }
            if (!chain_bad && curv != first_a) chain_bad = 1;
        }
        if (__ballot_sync(WFULL, chain_bad != 0)) return mi_warp_fail(cx, 2, lane);
        nf += cnt;
        nfe += tot;
    }
    const int n_faces_after_cut = nf;
    __syncwarp();

    // ---- rebuild
    // vertices: everything strictly inside M's region goes, except the tet corners
    int nv2 = 0;
    for (int base = 0; base < nv; base += 32) {
        const int v = base + lane;
        bool keep = false;
        if (v < nv) {
            const bool corner = cx.vm[v][2] < 4;
            keep = !(cx.vo[v] > 0 && !corner);
            if (keep && cx.vo[v] > 0) cx.vm[v][3] = (uint8_t)mid;
        }
        const unsigned mk = __ballot_sync(WFULL, keep);
        if (v < nv) cx.vmap[v] = keep ? (I)(nv2 + __popc(mk & lt)) : NI;
        nv2 += __popc(mk);
    }
    for (int e = lane; e < ne; e += 32) cx.merged_of[e] = NI;
    __syncwarp();
    // merged simplex-edge pieces: lane p < 6 owns simplex edge (i, j)
    int n_merged = 0;
    if (any_pos) {
        const uint8_t pi[6] = {0, 0, 0, 1, 1, 2}, pj[6] = {1, 2, 3, 2, 3, 3};
        int start = -1, end = -1, cnt = 0;
        const int i = pi[lane % 6], j = pj[lane % 6];
        auto member = [&](int e) {
            if (e < nE && cx.ec_split[e]) return false;
            return cx.em[e][1] < 4 && cx.em[e][0] == i && cx.em[e][1] == j && cx.edge_is_positive(e, nE);
        };
        if (lane < 6)
            for (int e = 0; e < ne; ++e) {
                if (!member(e)) continue;
                ++cnt;
                bool has_pred = false, has_succ = false;
                for (int g = 0; g < ne; ++g) {
                    if (!member(g)) continue;
                    if (cx.ev1[g] == cx.ev0[e]) has_pred = true;
                    if (cx.ev0[g] == cx.ev1[e]) has_succ = true;
                }
                if (!has_pred) start = cx.ev0[e];
                if (!has_succ) end = cx.ev1[e];
            }
        const unsigned mm = __ballot_sync(WFULL, lane < 6 && cnt > 0);
        n_merged = __popc(mm);
        if (lane < 6 && cnt > 0) {
            const int r = __popc(mm & lt);
            for (int e = 0; e < ne; ++e)
                if (member(e)) cx.merged_of[e] = (I)r;
            sc.mg_v0[r] = (I)start;
            sc.mg_v1[r] = (I)end;
            sc.mg_i[r] = (uint8_t)i;
            sc.mg_j[r] = (uint8_t)j;
        }
    }
    // edge remap (ec_neg reused): survivors in order, merged edges appended
    int n_surv = 0;
    for (int base = 0; base < ne; base += 32) {
        const int e = base + lane;
        bool alive = false;
        if (e < ne) alive = !((e < nE && cx.ec_split[e]) || (any_pos && cx.edge_is_positive(e, nE)));
        const unsigned ma = __ballot_sync(WFULL, alive);
        __syncwarp(); // edge_is_positive of the other lanes reads ec_pos only; ec_neg is ours to overwrite
        if (e < ne) cx.ec_neg[e] = alive ? (I)(n_surv + __popc(ma & lt)) : NI;
        n_surv += __popc(ma);
    }
    const int merged_base = n_surv;
    // cells: negative / untouched cells keep their order, then the negative parts of split cells, then M's cell
    int n_new_cells = 0;
    for (int base = 0; base < nC; base += 32) {
        const int c = base + lane;
        const bool k = (c < nC) && (cx.cstat[c] == C_NEG || cx.cstat[c] == C_ZERO);
        const unsigned mk = __ballot_sync(WFULL, k);
        if (c < nC) cx.cmap[c] = k ? (uint8_t)(n_new_cells + __popc(mk & lt)) : N8;
        n_new_cells += __popc(mk);
    }
    for (int base = 0; base < nC; base += 32) {
        const int c = base + lane;
        const bool k = (c < nC) && cx.cstat[c] == C_SPLIT;
        const unsigned mk = __ballot_sync(WFULL, k);
        if (c < nC) cx.cneg[c] = k ? (uint8_t)(n_new_cells + __popc(mk & lt)) : N8;
        n_new_cells += __popc(mk);
    }
    const int new_cell = any_pos ? n_new_cells++ : (int)N8;
    if (n_new_cells > Caps::MAXC) return mi_warp_fail(cx, 1, lane);
    __syncwarp();
    for (int c = lane; c < nC; c += 32) {
        if (cx.cmap[c] != N8) sc.new_cmat[cx.cmap[c]] = cx.cmat[c];
        if (cx.cneg[c] != N8) sc.new_cmat[cx.cneg[c]] = cx.cmat[c];
    }
    if (any_pos && lane == 0) sc.new_cmat[new_cell] = (uint8_t)mid;
    __syncwarp();
    auto map_edge = [&](int e) -> int {
        return cx.merged_of[e] != NI ? merged_base + (int)cx.merged_of[e] : (int)cx.ec_neg[e];
    };
    auto side_cell = [&](int old_cell) -> int {
        if (old_cell == N8) return N8;
        switch (cx.cstat[old_cell]) {
        case C_NEG:
        case C_ZERO: return cx.cmap[old_cell];
        case C_POS: return new_cell;
        default: return cx.cneg[old_cell];
        }
    };
    // faces -> buffer B2; positive boundary pieces are collected per simplex face instead
    int nf2 = 0, nfe2 = 0;
    int nbp[4] = {0, 0, 0, 0};
    for (int base = 0; base < n_faces_after_cut; base += 32) {
        const int f = base + lane;
        int kind = 0; // 1 emit, 2 boundary piece of M's region
        int pc = N8, ncell = N8, n = 0, off = 0, bface = 0;
        bool flip = false;
        if (f < n_faces_after_cut && !(f < nF && cx.fc_split[f])) {
            n = cx.flen[B][f];
            off = cx.foff[B][f];
            const bool is_cut_face = (f >= nF) && cx.fb[B][f] == N8 && cx.fpos[B][f] == N8 && cx.fneg[B][f] == N8;
            bool fpositive = false;
            if (f < nF)
                fpositive = (cx.fc_pos[f] == f);
            else if (!is_cut_face)
                for (int k = 0; k < n; ++k) fpositive |= (cx.vo[cx.fv[B][off + k]] > 0);
            if (is_cut_face) {
                pc = new_cell;
                ncell = cx.cneg[cx.fc_cut[f]];
                kind = 1;
            } else if (any_pos && fpositive) {
                if (cx.fb[B][f] != N8) {
                    kind = 2;
                    bface = cx.fb[B][f];
                }
            } else {
                pc = side_cell(cx.fpos[B][f]);
                ncell = side_cell(cx.fneg[B][f]);
                if (cx.fb[B][f] == N8 && pc != N8 && ncell != N8 && sc.new_cmat[pc] < sc.new_cmat[ncell]) {
                    const int t = pc;
                    pc = ncell;
                    ncell = t;
                    flip = true;
                }
                kind = 1;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned mb = __ballot_sync(WFULL, kind == 2 && bface == i);
            if (!mb) continue;
            if (nbp[i] + __popc(mb) > 12) return mi_warp_fail(cx, 1, lane);
            if (kind == 2 && bface == i) sc.bp[i][nbp[i] + __popc(mb & lt)] = (I)f;
            nbp[i] += __popc(mb);
        }
        const unsigned me = __ballot_sync(WFULL, kind == 1);
        if (!me) continue;
        int tot;
        const int po = nfe2 + warp_excl_scan(kind == 1 ? n : 0, lane, tot);
        const int cnt = __popc(me);
        if (nf2 + cnt > Caps::MAXF || nfe2 + tot > Caps::MAXFE) return mi_warp_fail(cx, 1, lane);
        if (kind == 1) {
            const int d = nf2 + __popc(me & lt);
            cx.foff[B2][d] = (uint16_t)po;
            cx.flen[B2][d] = (uint8_t)n;
            cx.fb[B2][d] = cx.fb[B][f];
            cx.fpos[B2][d] = (uint8_t)pc;
            cx.fneg[B2][d] = (uint8_t)ncell;
            for (int k = 0; k < n; ++k) {
                const int sv = flip ? (n - 1 - k) : k;
                const int se = flip ? ((2 * n - 2 - k) % n) : k;
                cx.fv[B2][po + k] = cx.vmap[cx.fv[B][off + sv]];
                cx.fe[B2][po + k] = (I)map_edge(cx.fe[B][off + se]);
            }
        }
        nf2 += cnt;
        nfe2 += tot;
    }
    __syncwarp();
    // merged boundary faces: lane i < 4 owns simplex face i
    {
        I sg_from[Caps::MAXLOOP], sg_to[Caps::MAXLOOP], sg_e[Caps::MAXLOOP];
        int ns = 0, lerr = 0;
        const bool mine = lane < 4 && nbp[lane & 3] > 0;
        if (mine) {
            const int i = lane;
            for (int q = 0; q < nbp[i] && !lerr; ++q) {
                const int f = sc.bp[i][q], n = cx.flen[B][f], off = cx.foff[B][f];
                for (int k = 0; k < n; ++k) {
                    const int e = cx.fe[B][off + k];
                    const int nee = map_edge(e);
                    if (nee == NI) continue;
                    bool dup = false;
                    for (int s = 0; s < ns; ++s) dup |= (sg_e[s] == nee);
                    if (dup) continue;
                    int a0, a1; // end points of the new edge (old vertex numbering)
                    if (cx.merged_of[e] != NI) {
                        a0 = sc.mg_v0[cx.merged_of[e]];
                        a1 = sc.mg_v1[cx.merged_of[e]];
                    } else {
                        a0 = cx.ev0[e];
                        a1 = cx.ev1[e];
                    }
                    const bool forward = (cx.ev0[e] == cx.fv[B][off + k]);
                    if (ns >= Caps::MAXLOOP) {
                        lerr = 1;
                        break;
                    }
                    sg_from[ns] = cx.vmap[forward ? a0 : a1];
                    sg_to[ns] = cx.vmap[forward ? a1 : a0];
                    sg_e[ns] = (I)nee;
                    ++ns;
                }
            }
        }
        const unsigned mm = __ballot_sync(WFULL, mine);
        if (mm) {
            if (__ballot_sync(WFULL, lerr != 0)) return mi_warp_fail(cx, 1, lane);
            int tot;
            const int po = nfe2 + warp_excl_scan(mine ? ns : 0, lane, tot);
            const int cnt = __popc(mm);
            if (nf2 + cnt > Caps::MAXF || nfe2 + tot > Caps::MAXFE) return mi_warp_fail(cx, 1, lane);
            int chain_bad = 0;
            if (mine) {
                const int d = nf2 + __popc(mm & lt);
                cx.foff[B2][d] = (uint16_t)po;
                cx.flen[B2][d] = (uint8_t)ns;
                cx.fb[B2][d] = (uint8_t)lane;
                cx.fpos[B2][d] = N8;
                cx.fneg[B2][d] = (uint8_t)new_cell;
                unsigned long long used = 0;
                int curv = sg_from[0], w = po;
                for (int step = 0; step < ns; ++step) {
                    int pick = -1;
                    for (int k = 0; k < ns; ++k)
                        if (!((used >> k) & 1) && sg_from[k] == curv) {
                            pick = k;
                            break;
                        }
                    if (pick < 0) {
                        chain_bad = 1;
                        break;
                    }
                    used |= 1ull << pick;
                    cx.fv[B2][w] = (I)curv;
                    cx.fe[B2][w] = sg_e[pick];
                    ++w;
                    curv = sg_to[pick];
                }
                if (!chain_bad && curv != sg_from[0]) chain_bad = 1;
            }
            if (__ballot_sync(WFULL, chain_bad != 0)) return mi_warp_fail(cx, 2, lane);
            nf2 += cnt;
            nfe2 += tot;
        }
    }
    __syncwarp();
    // ---- commit: edges (in place; merged appended), vertices, cells, faces
    for (int base = 0; base < ne; base += 32) {
        const int e = base + lane;
        int d = NI;
        I a0 = 0, a1 = 0;
        uint8_t m0 = 0, m1 = 0, m2 = 0;
        if (e < ne) {
            d = cx.ec_neg[e];
            if (d != NI) {
                a0 = cx.vmap[cx.ev0[e]];
                a1 = cx.vmap[cx.ev1[e]];
                m0 = cx.em[e][0];
                m1 = cx.em[e][1];
                m2 = cx.em[e][2];
            }
        }
        __syncwarp();
        if (d != NI) {
            cx.ev0[d] = a0;
            cx.ev1[d] = a1;
            cx.em[d][0] = m0;
            cx.em[d][1] = m1;
            cx.em[d][2] = m2;
        }
        __syncwarp();
    }
    if (merged_base + n_merged > Caps::MAXE) return mi_warp_fail(cx, 1, lane);
    if (lane < n_merged) {
        const int g = lane;
        cx.ev0[merged_base + g] = cx.vmap[sc.mg_v0[g]];
        cx.ev1[merged_base + g] = cx.vmap[sc.mg_v1[g]];
        cx.em[merged_base + g][0] = sc.mg_i[g];
        cx.em[merged_base + g][1] = sc.mg_j[g];
        cx.em[merged_base + g][2] = (uint8_t)mid;
    }
    __syncwarp();
    for (int base = 0; base < nv; base += 32) {
        const int v = base + lane;
        int d = NI;
        uint8_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        if (v < nv) {
            d = cx.vmap[v];
            if (d != NI) {
                q0 = cx.vm[v][0];
                q1 = cx.vm[v][1];
                q2 = cx.vm[v][2];
                q3 = cx.vm[v][3];
            }
        }
        __syncwarp();
        if (d != NI) {
            cx.vm[d][0] = q0;
            cx.vm[d][1] = q1;
            cx.vm[d][2] = q2;
            cx.vm[d][3] = q3;
        }
        __syncwarp();
    }
    for (int c = lane; c < n_new_cells; c += 32) cx.cmat[c] = sc.new_cmat[c];
    if (lane == 0) {
        cx.ne = merged_base + n_merged;
        cx.nv = nv2;
        cx.nc = n_new_cells;
        cx.nf = nf2;
        cx.nfe = nfe2;
        cx.cur = B2;
    }
    __syncwarp();
    return duplicate_of;
}

template <class Caps>
__device__ void warp_insert_material(MIComplex<Caps>& cx, MIWarpScratch<Caps>& sc, const double v[4], int lane)
{
    __syncwarp();
    const int err0 = cx.err, mid = cx.nm;
    __syncwarp(); // every lane has read the state before lane 0 may change it
    if (err0) return;
    if (mid >= Caps::MAXK + 4) {
        mi_warp_fail(cx, 1, lane);
        return;
    }
    if (lane < 4) cx.mval[mid - 4][lane] = v[lane];
    if (lane == 0) cx.nm = mid + 1;
    __syncwarp();
    const int dup = warp_add_material(cx, sc, mid, lane);
    if (cx.err) return;
    if (lane == 0) {
        if (dup < 0)
            cx.umi[mid] = (uint8_t)cx.n_groups++;
        else {
            cx.umi[mid] = cx.umi[dup];
            cx.has_dup = true;
        }
    }
    __syncwarp();
}

} // namespace rin
