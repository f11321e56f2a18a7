// Warp-cooperative plane insertion: one tetrahedron per warp, the complex in shared memory, every
// phase of the insertion (vertex classification, edge / face / cell cuts, consolidation, iso record)
// data-parallel over the entities with ballot / shuffle prefix sums for the new ids.
//
// Same algorithm and the same numbering as IAComplex::add_plane (ia_complex.cuh), which replaces
// compute_arrangement() of the un-vendored simplicial_arrangement library
// (/root/reference/src/implicit_arrangement.cpp:279,283): entity e of chunk lane L gets the id the
// serial loop would have given it, because ids are "running count + rank among the lanes before me".
// Only the layout of the loop pools between insertions may differ (they are re-packed every time).
//
// All 32 lanes call every function with identical arguments; `lane` is threadIdx.x & 31.
// tests/simt/ runs this file on the CPU (32 threads per warp) against the serial version.
#pragma once
#include "ia_complex.cuh"

namespace rin {

constexpr unsigned WFULL = 0xffffffffu;


__device__ __forceinline__ int warp_excl_scan(int x, int lane, int& total)
{
    int inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(WFULL, inc, d);
        if (lane >= d) inc += y;
    }
    total = __shfl_sync(WFULL, inc, 31);
    return inc - x;
}

// scratch pools for the re-packing step (shared memory, next to the complex)
template <class Caps>
struct IAWarpScratch
{
    using I = typename Caps::idx;
    I fv2[Caps::MAXFE], fe2[Caps::MAXFE];
    I cf2[Caps::MAXCF];
};

// uniform error exit
template <class Caps>
__device__ __forceinline__ int warp_fail(IAComplex<Caps>& cx, int code, int lane)
{
    if (lane == 0) cx.err = code;
    __syncwarp();
    return -1;
}

// insert plane `pid` (values already in plv); returns a coincident plane id or -1
template <class Caps>
__device__ int warp_add_plane(IAComplex<Caps>& cx, IAWarpScratch<Caps>& sc, int pid, int lane)
{
    using I = typename Caps::idx;
    constexpr I NI = IAComplex<Caps>::NI;
    const unsigned lt = (1u << lane) - 1u;
    const double* q = cx.plv[pid - 4];
    int nv = cx.nv, ne = cx.ne, nf = cx.nf, nc = cx.nc, nfe = cx.nfe, ncf = cx.ncf;

    // ---- 1. vertices
    {
        unsigned nex = 0;
        int bad = 0;
        bool anyl = false;
        for (int v = lane; v < nv; v += 32) {
            const int o = cx.orient_vertex(v, q, nex, bad);
            cx.vo[v] = (int8_t)o;
            anyl |= (o != 0);
        }
        const bool any = __ballot_sync(WFULL, anyl) != 0u;
        const bool degenerate = __ballot_sync(WFULL, bad != 0) != 0u;
        if (__ballot_sync(WFULL, nex != 0u)) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) nex += __shfl_xor_sync(WFULL, nex, d);
            if (lane == 0) cx.n_exact += nex;
        }
        if (!any || degenerate) return warp_fail(cx, 2, lane);
        __syncwarp();
    }

    // ---- 2. edges
    const int nE = ne;
    for (int base = 0; base < nE; base += 32) {
        const int e = base + lane;
        bool split = false;
        int o0 = 0;
        if (e < nE) {
            o0 = cx.vo[cx.ev0[e]];
            const int o1 = cx.vo[cx.ev1[e]];
            I p = NI, n = NI, x = NI;
            if (!(o0 == 0 && o1 == 0)) {
                if (o0 == 0)
                    x = cx.ev0[e];
                else if (o1 == 0)
                    x = cx.ev1[e];
                if (o0 >= 0 && o1 >= 0)
                    p = (I)e;
                else if (o0 <= 0 && o1 <= 0)
                    n = (I)e;
                else
                    split = true;
            }
            cx.ec_pos[e] = p;
            cx.ec_neg[e] = n;
            cx.ec_x[e] = x;
            cx.ec_split[e] = split ? 1 : 0;
        }
        const unsigned ms = __ballot_sync(WFULL, split);
        if (!ms) continue;
        const int cnt = __popc(ms);
        if (nv + cnt > Caps::MAXV || ne + 2 * cnt > Caps::MAXE) return warp_fail(cx, 1, lane);
        if (split) {
            const int r = __popc(ms & lt);
            const int x = nv + r, a = ne + 2 * r, b = a + 1;
            cx.vp[x][0] = cx.ep0[e];
            cx.vp[x][1] = cx.ep1[e];
            cx.vp[x][2] = (uint8_t)pid;
            cx.vo[x] = 0;
            cx.ec_x[e] = (I)x;
            cx.ec_pos[e] = (I)a; // sub-edges keep the direction v0 -> x -> v1; the positive one first
            cx.ec_neg[e] = (I)b;
            const int first = (o0 > 0) ? a : b, second = (o0 > 0) ? b : a;
            cx.ev0[first] = cx.ev0[e];
            cx.ev1[first] = (I)x;
            cx.ev0[second] = (I)x;
            cx.ev1[second] = cx.ev1[e];
            cx.ep0[a] = cx.ep0[b] = cx.ep0[e];
            cx.ep1[a] = cx.ep1[b] = cx.ep1[e];
            cx.ec_split[a] = cx.ec_split[b] = 0;
        }
        nv += cnt;
        ne += 2 * cnt;
    }
    __syncwarp();

    // ---- 3. faces
    const int nF = nf;
    int coplanar = -1;
    for (int base = 0; base < nF; base += 32) {
        const int f = base + lane;
        int kind = 0; // 1 on the plane, 2 positive, 3 negative, 4 split
        int n = 0, off = 0;
        if (f < nF) {
            n = cx.flen[f];
            off = cx.foff[f];
            int npos = 0, nneg = 0;
            for (int k = 0; k < n; ++k) {
                const int o = cx.vo[cx.fv[off + k]];
                npos += (o > 0);
                nneg += (o < 0);
            }
            kind = (npos == 0 && nneg == 0) ? 1 : (nneg == 0 ? 2 : (npos == 0 ? 3 : 4));
            cx.fc_pos[f] = (kind == 2) ? (I)f : NI;
            cx.fc_neg[f] = (kind == 3) ? (I)f : NI;
            cx.fc_cut[f] = NI;
            cx.fc_split[f] = (kind == 4) ? 1 : 0;
        }
        const unsigned mc = __ballot_sync(WFULL, kind == 1);
        if (mc && coplanar < 0) {
            const int pl = (kind == 1) ? (int)cx.fplane[f] : 0;
            coplanar = __shfl_sync(WFULL, pl, __ffs(mc) - 1);
        }
        const unsigned ms = __ballot_sync(WFULL, kind == 4);
        if (!ms) continue;
#define RIN_O(k) ((int)cx.vo[cx.fv[off + ((k) % n)]])
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
            return warp_fail(cx, 1, lane);
        if (kind == 4) {
            const int r = __popc(ms & lt);
            const int ei = cx.fe[off + (i % n)], ejl = cx.fe[off + (jl % n)];
            int start_tv, end_tv, first_pos, last_pos, first_neg, last_neg;
            if (RIN_O(i) == 0) {
                start_tv = cx.fv[off + (i % n)];
                first_pos = ei;
                last_neg = cx.fe[off + ((i + n - 1) % n)];
            } else {
                start_tv = cx.ec_x[ei];
                first_pos = cx.ec_pos[ei];
                last_neg = cx.ec_neg[ei];
            }
            if (RIN_O(jl + 1) == 0) {
                end_tv = cx.fv[off + ((jl + 1) % n)];
                last_pos = ejl;
                first_neg = cx.fe[off + ((jl + 1) % n)];
            } else {
                end_tv = cx.ec_x[ejl];
                last_pos = cx.ec_pos[ejl];
                first_neg = cx.ec_neg[ejl];
            }
            const int ce = ne + r;
            cx.ev0[ce] = (I)start_tv;
            cx.ev1[ce] = (I)end_tv;
            cx.ep0[ce] = cx.fplane[f];
            cx.ep1[ce] = (uint8_t)pid;
            cx.ec_split[ce] = 0;
            cx.fc_cut[f] = (I)ce;
            const int P = nf + 2 * r, Ng = P + 1;
            cx.fc_pos[f] = (I)P;
            cx.fc_neg[f] = (I)Ng;
            cx.fplane[P] = cx.fplane[Ng] = cx.fplane[f];
            cx.fpos[P] = cx.fpos[Ng] = cx.fpos[f];
            cx.fneg[P] = cx.fneg[Ng] = cx.fneg[f];
            cx.fc_split[P] = cx.fc_split[Ng] = 0;
            int w = nfe + poff;
            // positive loop
            cx.foff[P] = (uint16_t)w;
            cx.fv[w] = (I)start_tv;
            cx.fe[w] = (I)first_pos;
            ++w;
            for (int k = i + 1; k <= jl; ++k) {
                cx.fv[w] = cx.fv[off + (k % n)];
                cx.fe[w] = (k == jl) ? (I)last_pos : cx.fe[off + (k % n)];
                ++w;
            }
            cx.fv[w] = (I)end_tv;
            cx.fe[w] = (I)ce;
            ++w;
            cx.flen[P] = (uint8_t)(w - cx.foff[P]);
            // negative loop
            cx.foff[Ng] = (uint16_t)w;
            cx.fv[w] = (I)end_tv;
            cx.fe[w] = (I)first_neg;
            ++w;
            for (int k = kfirst; k <= klast; ++k) {
                cx.fv[w] = cx.fv[off + (k % n)];
                cx.fe[w] = (k == klast) ? (I)last_neg : cx.fe[off + (k % n)];
                ++w;
            }
            cx.fv[w] = (I)start_tv;
            cx.fe[w] = (I)ce;
            ++w;
            cx.flen[Ng] = (uint8_t)(w - cx.foff[Ng]);
        }
#undef RIN_O
        ne += cnt;
        nf += 2 * cnt;
        nfe += tot;
    }
    __syncwarp();

    // ---- 4. cells: phase A classifies and collects the cut polygon's edges, phase B writes
    const int nC = nc;
    for (int base = 0; base < nC; base += 32) {
        const int c = base + lane;
        bool split = false;
        int len_pos = 0, len_neg = 0, n_cut = 0, first_a = -1, first_b = -1, lerr = 0;
        int cl = 0, co = 0;
        I cut_e[Caps::MAXLOOP];
        if (c < nC) {
            cl = cx.clen[c];
            co = cx.coff[c];
            for (int k = 0; k < cl; ++k) {
                const int f = cx.cf[co + k];
                len_pos += (cx.fc_pos[f] != NI);
                len_neg += (cx.fc_neg[f] != NI);
            }
            split = len_pos > 0 && len_neg > 0;
            cx.c_split[c] = split ? 1 : 0;
            if (split) {
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
                for (int k = 0; k < cl; ++k) {
                    const int f = cx.cf[co + k];
                    const bool inward = (cx.fpos[f] == c);
                    if (cx.fc_split[f]) {
                        const int ce = cx.fc_cut[f];
                        add_cut_edge(ce, cx.ev0[ce], cx.ev1[ce], inward, true);
                    } else if (cx.fc_pos[f] != NI || cx.fc_neg[f] != NI) {
                        const int n = cx.flen[f], off = cx.foff[f];
                        for (int j = 0; j < n; ++j) {
                            const int a = cx.fv[off + j], b = cx.fv[off + ((j + 1) % n)];
                            if (cx.vo[a] == 0 && cx.vo[b] == 0)
                                add_cut_edge(cx.fe[off + j], a, b, inward, cx.fc_neg[f] != NI);
                        }
                    }
                }
            }
        }
        const unsigned ms = __ballot_sync(WFULL, split);
        if (!ms) continue;
        if (__ballot_sync(WFULL, lerr != 0)) return warp_fail(cx, 1, lane);
        int tot_cf, tot_fe;
        const int cfo = warp_excl_scan(split ? len_pos + len_neg + 2 : 0, lane, tot_cf);
        const int feo = warp_excl_scan(split ? n_cut : 0, lane, tot_fe);
        const int cnt = __popc(ms);
        if (nc + 2 * cnt > Caps::MAXC || nf + cnt > Caps::MAXF || ncf + tot_cf > Caps::MAXCF ||
            nfe + tot_fe > Caps::MAXFE)
            return warp_fail(cx, 1, lane);
        __syncwarp();
        int chain_bad = 0;
        if (split) {
            const int r = __popc(ms & lt);
            const int cp = nc + 2 * r, cn = cp + 1, G = nf + r;
            cx.c_split[cp] = cx.c_split[cn] = 0;
            int w = ncf + cfo;
            cx.coff[cp] = (uint16_t)w;
            for (int k = 0; k < cl; ++k) {
                const int f = cx.cf[co + k];
                if (cx.fc_pos[f] != NI) cx.cf[w++] = cx.fc_pos[f];
            }
            cx.cf[w++] = (I)G;
            cx.clen[cp] = (uint8_t)(w - cx.coff[cp]);
            cx.coff[cn] = (uint16_t)w;
            for (int k = 0; k < cl; ++k) {
                const int f = cx.cf[co + k];
                if (cx.fc_neg[f] != NI) cx.cf[w++] = cx.fc_neg[f];
            }
            cx.cf[w++] = (I)G;
            cx.clen[cn] = (uint8_t)(w - cx.coff[cn]);
            // the cut face G: chain the cut edges into a loop starting first_a -> first_b
            cx.fc_pos[G] = cx.fc_neg[G] = cx.fc_cut[G] = NI;
            cx.fc_split[G] = 0;
            cx.fplane[G] = (uint8_t)pid;
            cx.fpos[G] = (uint8_t)cp;
            cx.fneg[G] = (uint8_t)cn;
            cx.foff[G] = (uint16_t)(nfe + feo);
            cx.flen[G] = (uint8_t)n_cut;
            {
                int wf = nfe + feo;
                unsigned long long used = 0;
                int cur = first_a;
                for (int step = 0; step < n_cut && !chain_bad; ++step) {
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
                        } else if (cx.ev0[e] == cur || cx.ev1[e] == cur) {
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
                    cx.fv[wf] = (I)cur;
                    cx.fe[wf] = (I)e;
                    ++wf;
                    cur = (cx.ev0[e] == cur) ? cx.ev1[e] : cx.ev0[e];
                }
                if (cur != first_a) chain_bad = 1;
            }
            // which side of each sub-cell face is c: noted now (scratch pool, same positions as cf), written
            // after the barrier, because a face shared with another split cell is updated by that lane too
            for (int k = cx.coff[cp]; k < w; ++k) {
                const int f = cx.cf[k];
                sc.cf2[k] = (f != G && cx.fpos[f] == c) ? 1 : 0;
            }
        }
        if (__ballot_sync(WFULL, chain_bad != 0)) return warp_fail(cx, 2, lane);
        __syncwarp();
        if (split) {
            // the sub-cells' faces now border cp / cn instead of c (each face has c on one side only)
            const int r = __popc(ms & lt);
            const int cp = nc + 2 * r, cn = cp + 1;
            for (int half = 0; half < 2; ++half) {
                const int d = half ? cn : cp;
                for (int k = 0; k < cx.clen[d] - 1; ++k) {
                    const int pos = cx.coff[d] + k, f = cx.cf[pos];
                    if (sc.cf2[pos])
                        cx.fpos[f] = (uint8_t)d;
                    else
                        cx.fneg[f] = (uint8_t)d;
                }
            }
        }
        nc += 2 * cnt;
        nf += cnt;
        ncf += tot_cf;
        nfe += tot_fe;
        __syncwarp();
    }
    __syncwarp();

    // ---- 5. re-pack (survivors keep their order)
    {
        int k0 = 0;
        for (int base = 0; base < ne; base += 32) { // edges; ec_pos becomes the edge remap
            const int e = base + lane;
            const bool in = e < ne;
            const bool alive = in && !((e < nE) && cx.ec_split[e]);
            const unsigned ma = __ballot_sync(WFULL, alive);
            const int k = k0 + __popc(ma & lt);
            I a0 = 0, a1 = 0;
            uint8_t p0 = 0, p1 = 0;
            if (alive) {
                a0 = cx.ev0[e];
                a1 = cx.ev1[e];
                p0 = cx.ep0[e];
                p1 = cx.ep1[e];
            }
            __syncwarp();
            if (in) cx.ec_pos[e] = alive ? (I)k : NI;
            if (alive) {
                cx.ev0[k] = a0;
                cx.ev1[k] = a1;
                cx.ep0[k] = p0;
                cx.ep1[k] = p1;
            }
            k0 += __popc(ma);
            __syncwarp();
        }
        ne = k0;
        int kc = 0;
        for (int base = 0; base < nc; base += 32) { // cell remap
            const int c = base + lane;
            const bool alive = (c < nc) && !((c < nC) && cx.c_split[c]);
            const unsigned ma = __ballot_sync(WFULL, alive);
            if (c < nc) cx.cmap[c] = alive ? (uint8_t)(kc + __popc(ma & lt)) : N8;
            kc += __popc(ma);
        }
        __syncwarp();
        int kf = 0, pool = 0;
        for (int base = 0; base < nf; base += 32) { // faces; fc_pos becomes the face remap
            const int f = base + lane;
            const bool in = f < nf;
            const bool alive = in && !((f < nF) && cx.fc_split[f]);
            const unsigned ma = __ballot_sync(WFULL, alive);
            const int k = kf + __popc(ma & lt);
            int n = 0, off = 0;
            uint8_t pl = 0, cp = N8, cn = N8;
            if (alive) {
                n = cx.flen[f];
                off = cx.foff[f];
                pl = cx.fplane[f];
                cp = cx.fpos[f];
                cn = cx.fneg[f];
            }
            int tot;
            const int po = pool + warp_excl_scan(n, lane, tot);
            for (int j = 0; j < n; ++j) {
                sc.fv2[po + j] = cx.fv[off + j];
                sc.fe2[po + j] = cx.ec_pos[cx.fe[off + j]];
            }
            __syncwarp();
            if (in) cx.fc_pos[f] = alive ? (I)k : NI;
            if (alive) {
                cx.foff[k] = (uint16_t)po;
                cx.flen[k] = (uint8_t)n;
                cx.fplane[k] = pl;
                cx.fpos[k] = (cp == N8) ? N8 : cx.cmap[cp];
                cx.fneg[k] = (cn == N8) ? N8 : cx.cmap[cn];
            }
            kf += __popc(ma);
            pool += tot;
            __syncwarp();
        }
        nf = kf;
        nfe = pool;
        for (int p = lane; p < pool; p += 32) {
            cx.fv[p] = sc.fv2[p];
            cx.fe[p] = sc.fe2[p];
        }
        int cpool = 0;
        for (int base = 0; base < nc; base += 32) { // cells
            const int c = base + lane;
            const bool alive = (c < nc) && cx.cmap[c] != N8;
            int n = 0, off = 0, d = 0;
            if (alive) {
                n = cx.clen[c];
                off = cx.coff[c];
                d = cx.cmap[c];
            }
            int tot;
            const int po = cpool + warp_excl_scan(n, lane, tot);
            for (int j = 0; j < n; ++j) sc.cf2[po + j] = cx.fc_pos[cx.cf[off + j]];
            __syncwarp();
            if (alive) {
                cx.coff[d] = (uint16_t)po;
                cx.clen[d] = (uint8_t)n;
            }
            cpool += tot;
            __syncwarp();
        }
        for (int p = lane; p < cpool; p += 32) cx.cf[p] = sc.cf2[p];
        nc = kc;
        ncf = cpool;
    }
    if (lane == 0) {
        cx.nv = nv;
        cx.ne = ne;
        cx.nf = nf;
        cx.nc = nc;
        cx.nfe = nfe;
        cx.ncf = ncf;
    }
    __syncwarp();
    return coplanar;
}

// add the next input plane (values at the 4 corners); tracks coincident planes
template <class Caps>
__device__ void warp_insert(IAComplex<Caps>& cx, IAWarpScratch<Caps>& sc, const double v[4], int lane)
{
    __syncwarp();
    const int err0 = cx.err, pid = cx.np;
    __syncwarp(); // every lane has read the state before lane 0 may change it
    if (err0) return;
    if (pid >= Caps::MAXK + 4) {
        warp_fail(cx, 1, lane);
        return;
    }
    if (lane < 4) cx.plv[pid - 4][lane] = v[lane];
    if (lane == 0) cx.np = pid + 1;
    __syncwarp();
    const int cop = warp_add_plane(cx, sc, pid, lane);
    if (cx.err) return;
    if (lane == 0) {
        if (cop < 0)
            cx.upi[pid] = (uint8_t)cx.n_groups++;
        else {
            cx.upi[pid] = cx.upi[cop];
            cx.has_coplanar = true;
        }
    }
    __syncwarp();
}

} // namespace rin

