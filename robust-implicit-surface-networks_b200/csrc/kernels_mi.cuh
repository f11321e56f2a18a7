// sm_100a kernels of the material-interface hot path (reference: src/material_interface.cpp:53-447,
// src/extract_mesh.cpp:569-986, :1541-1637).  Shares the scan / dedup / face kernels with the
// implicit-arrangement path (kernels_ia.cuh).
#pragma once
#include "kernels_ia.cuh"
#include "mi_complex.cuh"

namespace rin {

// ---------------------------------------------------------------------------------------------
// K1 (MI): per-vertex set of materials attaining the maximum ("highest func" loop,
// src/material_interface.cpp:59-92) as a W-word bit mask in vmask[].x, and in vmask[].y bit 0 a
// flag "two materials (any two) are exactly equal here" used to detect materials tying on a
// whole tet face.  Runs after the values are in SoA form.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) highest_material_kernel(const double* __restrict__ vals,
    uint32_t v_first, uint32_t v_count, uint32_t V, uint32_t F, uint2* __restrict__ vmask,
    unsigned long long* __restrict__ n_tied)
{
    unsigned tied = 0;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < v_count; idx += gridDim.x * blockDim.x) {
        const uint32_t v = v_first + idx;
        double mx = vals[v];
        for (uint32_t f = 1; f < F; ++f) mx = fmax(mx, vals[(size_t)f * V + v]);
        unsigned cnt = 0;
        bool any_equal = false;
        if (F <= 32) {
            double x[32];
            for (uint32_t f = 0; f < F; ++f) x[f] = vals[(size_t)f * V + v];
            for (uint32_t f = 0; f < F; ++f)
                for (uint32_t g = f + 1; g < F; ++g) any_equal |= (x[f] == x[g]);
        } else {
            for (uint32_t f = 0; f < F && !any_equal; ++f) {
                const double xf = vals[(size_t)f * V + v];
                for (uint32_t g = f + 1; g < F && !any_equal; ++g) any_equal = (vals[(size_t)g * V + v] == xf);
            }
        }
        for (uint32_t w = 0; w * 32 < F; ++w) {
            uint32_t H = 0;
            const uint32_t fe = min(F, w * 32 + 32);
            for (uint32_t f = w * 32; f < fe; ++f)
                if (vals[(size_t)f * V + v] == mx) {
                    H |= 1u << (f & 31);
                    ++cnt;
                }
            vmask[(size_t)w * V + v] = make_uint2(H, (w == 0 && any_equal) ? 1u : 0u);
        }
        tied += (cnt > 1);
    }
    for (int o = 16; o; o >>= 1) tied += __shfl_xor_sync(0xffffffffu, tied, o);
    if ((threadIdx.x & 31) == 0 && tied) atomicAdd(n_tied, (unsigned long long)tied);
}

// ---------------------------------------------------------------------------------------------
// K2b: material filter + tile-local compaction (the "filter" loop, src/material_interface.cpp:99-152).
// S = union of the vertices' maximal materials; fewer than two -> no interface.  Otherwise
// min_h[c] = min over S of the value at corner c and every material exceeding min_h at >= 2
// corners joins S.  The cheap reject needs only the 4 gathered masks; the value test runs for the
// few candidate tets.
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(FILT_THREADS) filter_mi_tiles_kernel(const uint4* __restrict__ tets,
    uint32_t t_first, uint32_t t_count, const uint2* __restrict__ vmask, const double* __restrict__ vals,
    uint32_t V, uint32_t F, uint32_t* __restrict__ tl_tet, uint32_t* __restrict__ tl_mask, size_t tl_stride,
    uint2* __restrict__ tile_cnt, FilterCounters* __restrict__ ctr, unsigned* __restrict__ n_tie_faces)
{
    __shared__ unsigned s_cnt[FILT_ITEMS][FILT_THREADS / 32];
    __shared__ unsigned s_kf[FILT_THREADS / 32];
    __shared__ unsigned s_k[3];
    const unsigned tile = blockIdx.x;
    const uint32_t base = tile * FILT_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3) s_k[threadIdx.x] = 0;
    uint32_t m[FILT_ITEMS][W];
    unsigned ball[FILT_ITEMS];
    unsigned k1 = 0, k2 = 0, km = 0, kf = 0, ties = 0;
#pragma unroll
    for (int j = 0; j < FILT_ITEMS; ++j) {
        const uint32_t i = base + j * FILT_THREADS + threadIdx.x;
        const uint4 tv = __ldg(&tets[t_first + min(i, t_count - 1)]);
        const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
        int kq = 0, ns = 0;
        unsigned tie_cnt = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            uint32_t s = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint2 g = __ldg(&vmask[(size_t)w * V + vv[c]]);
                s |= g.x;
                if (w == 0) tie_cnt += g.y & 1u;
            }
            m[j][w] = s;
            ns += __popc(s);
        }
        if (i >= t_count) ns = 0;
        if (ns >= 2) {
            double min_h[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) min_h[c] = 1.7976931348623157e308;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t s = m[j][w];
                while (s) {
                    const int f = w * 32 + __ffs(s) - 1;
                    s &= s - 1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) min_h[c] = fmin(min_h[c], __ldg(&vals[(size_t)f * V + vv[c]]));
                }
            }
#pragma unroll
            for (int w = 0; w < W; ++w)
                for (uint32_t f = w * 32; f < F && f < (uint32_t)w * 32 + 32; ++f) {
                    int greater = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) greater += (__ldg(&vals[(size_t)f * V + vv[c]]) > min_h[c]);
                    if (greater > 1) m[j][w] |= 1u << (f & 31);
                }
#pragma unroll
            for (int w = 0; w < W; ++w) kq += __popc(m[j][w]);
            if (tie_cnt >= 3) {
                // exact test: an active material g and any other material coincide at the three
                // corners of one tet face -> a boundary face piece may be a material interface
                bool found = false;
                for (uint32_t g = 0; g < F && !found; ++g) {
                    bool active = false;
#pragma unroll
                    for (int w = 0; w < W; ++w)
                        if ((g >> 5) == (uint32_t)w) active = (m[j][w] >> (g & 31)) & 1;
                    if (!active) continue;
                    double xg[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) xg[c] = __ldg(&vals[(size_t)g * V + vv[c]]);
                    for (uint32_t f = 0; f < F && !found; ++f) {
                        if (f == g) continue;
                        int eq = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) eq += (__ldg(&vals[(size_t)f * V + vv[c]]) == xg[c]);
                        found = (eq >= 3);
                    }
                }
                ties += found;
            }
        } else {
#pragma unroll
            for (int w = 0; w < W; ++w) m[j][w] = 0;
        }
        k1 += (kq == 2);
        k2 += (kq == 3);
        km += (kq > 3);
        kf += kq;
        ball[j] = __ballot_sync(0xffffffffu, kq > 0);
        if (lane == 0) s_cnt[j][warp] = __popc(ball[j]);
    }
    for (int o = 16; o; o >>= 1) {
        k1 += __shfl_xor_sync(0xffffffffu, k1, o);
        k2 += __shfl_xor_sync(0xffffffffu, k2, o);
        km += __shfl_xor_sync(0xffffffffu, km, o);
        kf += __shfl_xor_sync(0xffffffffu, kf, o);
        ties += __shfl_xor_sync(0xffffffffu, ties, o);
    }
    if (lane == 0) s_kf[warp] = kf;
    __syncthreads();
    if (lane == 0) {
        if (k1) atomicAdd(&s_k[0], k1);
        if (k2) atomicAdd(&s_k[1], k2);
        if (km) atomicAdd(&s_k[2], km);
        if (ties) atomicAdd(n_tie_faces, ties);
    }
    unsigned my_off[FILT_ITEMS];
    unsigned run = 0;
#pragma unroll
    for (int r = 0; r < FILT_ITEMS * (FILT_THREADS / 32) / 32; ++r) {
        const int e = r * 32 + lane;
        const unsigned c = s_cnt[e / (FILT_THREADS / 32)][e % (FILT_THREADS / 32)];
        unsigned x = c;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        const unsigned excl = run + x - c;
#pragma unroll
        for (int j = 0; j < FILT_ITEMS; ++j) {
            const int idx = j * (FILT_THREADS / 32) + warp;
            if (idx / 32 == r) my_off[j] = __shfl_sync(0xffffffffu, excl, idx % 32);
        }
        run += __shfl_sync(0xffffffffu, x, 31);
    }
    const size_t tbase = (size_t)tile * FILT_TILE;
#pragma unroll
    for (int j = 0; j < FILT_ITEMS; ++j) {
        if ((ball[j] >> lane) & 1) {
            const size_t pos = tbase + my_off[j] + __popc(ball[j] & ((1u << lane) - 1));
            tl_tet[pos] = t_first + base + j * FILT_THREADS + threadIdx.x;
#pragma unroll
            for (int w = 0; w < W; ++w) tl_mask[(size_t)w * tl_stride + pos] = m[j][w];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned kft = 0;
        for (int w = 0; w < FILT_THREADS / 32; ++w) kft += s_kf[w];
        tile_cnt[tile] = make_uint2(run, kft);
        if (s_k[0]) atomicAdd(&ctr->n_k1, s_k[0]);
        if (s_k[1]) atomicAdd(&ctr->n_k2, s_k[1]);
        if (s_k[2]) atomicAdd(&ctr->n_kmore, s_k[2]);
    }
}

// ---------------------------------------------------------------------------------------------
// MI iso record (32-bit words):
//   word 0            n_verts | n_faces << 8 | n_face_vertex_entries << 16
//   per vertex        { local vertex id }, { m0 | m1 << 8 | m2 << 16 | m3 << 24 }   (ascending)
//   per face          { local face id (16) | n << 24 }, { positive label | negative label << 8 },
//                     ceil(n/4) words of vertex ranks
// Only interface faces (positive label > 3) and their vertices are recorded
// (src/extract_mesh.cpp:641-651).
// ---------------------------------------------------------------------------------------------
template <class Caps>
struct MIIsoScan
{
    uint32_t isov[(Caps::MAXV + 31) / 32];
    int nvi, nfi, nfv, nfw;
    __device__ void run(const MIComplex<Caps>& cx)
    {
        const int B = cx.cur;
        for (int i = 0; i < (Caps::MAXV + 31) / 32; ++i) isov[i] = 0;
        nfi = nfv = nfw = 0;
        for (int f = 0; f < cx.nf; ++f)
            if (cx.is_mi_face(f)) {
                ++nfi;
                const int n = cx.flen[B][f];
                nfv += n;
                if (n > 127) nfv = 1 << 20;
                nfw += 1 + rec_face_words(n);
                for (int k = 0; k < n; ++k) {
                    int v = cx.fv[B][cx.foff[B][f] + k];
                    isov[v >> 5] |= 1u << (v & 31);
                }
            }
        nvi = 0;
        for (int i = 0; i < (Caps::MAXV + 31) / 32; ++i) nvi += __popc(isov[i]);
    }
    __device__ int rank(int v) const
    {
        int r = __popc(isov[v >> 5] & ((1u << (v & 31)) - 1u));
        for (int i = 0; i < (v >> 5); ++i) r += __popc(isov[i]);
        return r;
    }
    __device__ uint32_t size_bytes() const { return 4u * uint32_t(1 + 2 * nvi + nfw); }
    __device__ void write(const MIComplex<Caps>& cx, uint32_t* w) const
    {
        const int B = cx.cur;
        int p = 0;
        w[p++] = (uint32_t)nvi | ((uint32_t)nfi << 8) | ((uint32_t)nfv << 16);
        for (int v = 0; v < cx.nv; ++v)
            if ((isov[v >> 5] >> (v & 31)) & 1) {
                w[p++] = (uint32_t)v;
                w[p++] = (uint32_t)cx.vm[v][0] | ((uint32_t)cx.vm[v][1] << 8) | ((uint32_t)cx.vm[v][2] << 16) |
                         ((uint32_t)cx.vm[v][3] << 24);
            }
        for (int f = 0; f < cx.nf; ++f)
            if (cx.is_mi_face(f)) {
                const int n = cx.flen[B][f];
                w[p++] = (uint32_t)f | ((uint32_t)n << 24);
                w[p++] = (uint32_t)cx.pos_label(f) | ((uint32_t)cx.neg_label(f) << 8);
                for (int k0 = 0; k0 < n; k0 += 4) {
                    uint32_t x = 0;
                    for (int k = k0; k < n && k < k0 + 4; ++k)
                        x |= (uint32_t)rank(cx.fv[B][cx.foff[B][f] + k]) << (8 * (k - k0));
                    w[p++] = x;
                }
            }
    }
};

template <class Caps, int W>
__device__ bool general_mi_one(MIComplex<Caps>& cx, uint32_t a, const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc, bool last_tier)
{
    const uint4 tv = __ldg(&tets[act_tet[a]]);
    bool first = true;
    for (int w = 0; w < W; ++w) {
        uint32_t mm = act_mask[(size_t)w * cap + a];
        while (mm) {
            int f = w * 32 + __ffs(mm) - 1;
            mm &= mm - 1;
            double pv[4];
            pv[0] = __ldg(&vals[(size_t)f * V + tv.x]);
            pv[1] = __ldg(&vals[(size_t)f * V + tv.y]);
            pv[2] = __ldg(&vals[(size_t)f * V + tv.z]);
            pv[3] = __ldg(&vals[(size_t)f * V + tv.w]);
            if (first) {
                cx.init(pv);
                first = false;
            } else
                cx.insert(pv);
        }
    }
    MIIsoScan<Caps> iso;
    if (!cx.err) {
        iso.run(cx);
        if (iso.nvi > 255 || iso.nfi > 255 || iso.nfv > 65535) cx.err = 1;
    }
    if (cx.n_exact) atomicAdd(&gc->n_exact, cx.n_exact);
    if (cx.err == 1 && !last_tier) return false;
    if (cx.err) {
        if (atomicCAS(&gc->err, 0, cx.err == 1 ? RIN_ERR_CAPACITY : RIN_ERR_ARRANGEMENT) == 0)
            gc->err_tet = act_tet[a];
        rec_ref[a] = REF_GENERAL;
        return true;
    }
    const uint32_t szal = iso.size_bytes();
    const uint32_t off = atomicAdd(&gc->arena_top, szal);
    if (off + szal > arena_cap) {
        gc->arena_overflow = 1;
        rec_ref[a] = REF_GENERAL;
        return true;
    }
    iso.write(cx, reinterpret_cast<uint32_t*>(arena + off));
    rec_ref[a] = REF_GENERAL | (off >> 2);
    return true;
}

template <int W>
__global__ void __launch_bounds__(GEN_SMALL_WARPS * 32) general_mi_small_kernel(
    const uint4* __restrict__ tets, const uint32_t* __restrict__ act_tet,
    const uint32_t* __restrict__ act_mask, uint32_t cap, const uint32_t* __restrict__ small_list,
    uint32_t* __restrict__ ovf_list, const double* __restrict__ vals, uint32_t V,
    uint8_t* __restrict__ arena, uint32_t arena_cap, uint32_t* __restrict__ rec_ref,
    GeneralCounters* __restrict__ gc)
{
    extern __shared__ __align__(16) uint8_t s_raw_mi[];
    MIComplex<MICapsSmall>* s_cx = reinterpret_cast<MIComplex<MICapsSmall>*>(s_raw_mi);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane != 0) return;
    const uint32_t n = gc->n_small;
    MIComplex<MICapsSmall>& cx = s_cx[warp];
    for (uint32_t g = blockIdx.x * GEN_SMALL_WARPS + warp; g < n; g += gridDim.x * GEN_SMALL_WARPS) {
        const uint32_t a = small_list[g];
        if (!general_mi_one<MICapsSmall, W>(cx, a, tets, act_tet, act_mask, cap, vals, V, arena, arena_cap,
                rec_ref, gc, false))
            ovf_list[atomicAdd(&gc->n_ovf, 1u)] = a;
    }
}

template <int W>
__global__ void __launch_bounds__(GEN_THREADS) general_mi_big_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ ovf_list,
    const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc)
{
    const uint32_t nb = gc->n_big, n = nb + gc->n_ovf;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        MIComplex<MICaps> cx;
        const uint32_t a = g < nb ? big_list[g] : ovf_list[g - nb];
        general_mi_one<MICaps, W>(cx, a, tets, act_tet, act_mask, cap, vals, V, arena, arena_cap, rec_ref, gc,
            true);
    }
}

// ---------------------------------------------------------------------------------------------
// K3 (MI): two materials without ties go to the 16-entry table (sign pattern of m0 - m1 at the
// corners); everything else to the general kernels.  Dispatch of src/material_interface.cpp:320-328
// (the 3-material "secondary" table is served by the general kernel: same results).
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) classify_mi_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    uint32_t n_active, const double* __restrict__ vals, uint32_t V, const uint16_t* __restrict__ lut2,
    int use_lookup, uint32_t* __restrict__ rec_ref, uint32_t* __restrict__ small_list,
    uint32_t* __restrict__ big_list, GeneralCounters* __restrict__ gc)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        uint32_t m[W];
        int k = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            m[w] = act_mask[(size_t)w * cap + a];
            k += __popc(m[w]);
        }
        uint32_t ref = REF_GENERAL;
        if (use_lookup && k == 2) {
            const uint4 tv = __ldg(&tets[act_tet[a]]);
            const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
            const int f0 = nth_set_bit(m, W, 0), f1 = nth_set_bit(m, W, 1);
            int key = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double x = __ldg(&vals[(size_t)f0 * V + vv[c]]), y = __ldg(&vals[(size_t)f1 * V + vv[c]]);
                if (x == y) key = -1000;
                key |= (x > y ? 1 : 0) << c;
            }
            if (key >= 0) ref = lut2[key];
        }
        if (ref & REF_GENERAL) {
            atomicAdd(&gc->n_general, 1u);
            if (k <= MICapsSmall::MAXK)
                small_list[atomicAdd(&gc->n_small, 1u)] = a;
            else
                big_list[atomicAdd(&gc->n_big, 1u)] = a;
        }
        rec_ref[a] = ref;
    }
}

// ---------------------------------------------------------------------------------------------
// K5b (MI): vertex candidates + face records (extract_MI_mesh, src/extract_mesh.cpp:652-832).
// cand_key = (v0, v1, v2, m0 | m1 << 10 | m2 << 20) sorted corners and sorted material ids;
// cand_pay = (tet, local | size << 8 | dedup << 16, m0 | m1 << 16, m2 | m3 << 16).
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) emit_mi_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    uint32_t n_active, const uint32_t* __restrict__ rec_ref, const uint4* __restrict__ offs,
    const uint8_t* __restrict__ lut_blob, const uint8_t* __restrict__ arena, uint4* __restrict__ cand_key,
    uint4* __restrict__ cand_pay, uint4* __restrict__ face_hdr, uint32_t* __restrict__ fv_ref)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const uint32_t ref = rec_ref[a];
        const uint32_t* r = (ref & REF_GENERAL)
                                ? reinterpret_cast<const uint32_t*>(arena) + (size_t)(ref & ~REF_GENERAL)
                                : reinterpret_cast<const uint32_t*>(lut_blob) + ref;
        const uint32_t hdr = r[0];
        const int nv = hdr & 255, nf = (hdr >> 8) & 255;
        if (nv == 0 && nf == 0) continue;
        const uint32_t t = act_tet[a];
        const uint4 tv4 = __ldg(&tets[t]);
        const uint32_t tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
        uint32_t m[W];
#pragma unroll
        for (int w = 0; w < W; ++w) m[w] = act_mask[(size_t)w * cap + a];
        const uint4 o = offs[a];
        const uint32_t* p = r + 1;
        for (int i = 0; i < nv; ++i, p += 2) {
            const int local = p[0] & 0xffff;
            const uint32_t e = p[1];
            const int ids[4] = {(int)(e & 255), (int)((e >> 8) & 255), (int)((e >> 16) & 255), (int)(e >> 24)};
            const int nb = (ids[0] < 4) + (ids[1] < 4) + (ids[2] < 4) + (ids[3] < 4);
            // real materials follow the boundary ones (ascending local id == ascending global id)
            uint32_t mg[4] = {0xffffu, 0xffffu, 0xffffu, 0xffffu};
            for (int q = nb; q < 4; ++q) mg[q - nb] = (uint32_t)nth_set_bit(m, W, ids[q] - 4);
            uint4 key, pay;
            pay.x = t;
            if (nb == 0) {
                key = make_uint4(tv[0], tv[1], tv[2], tv[3]);
                pay.y = (uint32_t)local | (4u << 8);
            } else {
                unsigned on_b = 0;
                for (int q = 0; q < nb; ++q) on_b |= 1u << ids[q];
                uint32_t c[3] = {NONE32, NONE32, NONE32};
                int ncn = 0;
                for (int q = 0; q < 4; ++q)
                    if (!((on_b >> q) & 1)) c[ncn++] = tv[q];
                if (ncn >= 2 && c[0] > c[1]) {
                    uint32_t s = c[0];
                    c[0] = c[1];
                    c[1] = s;
                }
                if (ncn == 3) {
                    if (c[1] > c[2]) {
                        uint32_t s = c[1];
                        c[1] = c[2];
                        c[2] = s;
                    }
                    if (c[0] > c[1]) {
                        uint32_t s = c[0];
                        c[0] = c[1];
                        c[1] = s;
                    }
                }
                // sorted material ids are part of the key (:713-724, :751-757); none for a tet vertex
                uint32_t kw = 0x3fffffffu;
                if (nb == 2)
                    kw = mg[0] | (mg[1] << 10) | (0x3ffu << 20);
                else if (nb == 1)
                    kw = mg[0] | (mg[1] << 10) | (mg[2] << 20);
                else
                    mg[0] = 0xffffu; // on a tet vertex: material_indices stay unset (:805-812)
                key = make_uint4(c[0], c[1], c[2], kw);
                pay.y = (uint32_t)local | ((uint32_t)ncn << 8) | (1u << 16);
            }
            pay.z = mg[0] | (mg[1] << 16);
            pay.w = mg[2] | (mg[3] << 16);
            cand_key[o.x + i] = key;
            cand_pay[o.x + i] = pay;
        }
        uint32_t fvo = o.z;
        for (int j = 0; j < nf; ++j) {
            const uint32_t e0 = *p++, e1 = *p++;
            const uint32_t local = e0 & 0xffffu;
            const int n = (e0 >> 24) & 127;
            const uint32_t fpos = (uint32_t)nth_set_bit(m, W, (int)(e1 & 255) - 4);
            const uint32_t fneg = (uint32_t)nth_set_bit(m, W, (int)((e1 >> 8) & 255) - 4);
            face_hdr[o.y + j] = make_uint4(t, local | ((uint32_t)n << 16), fpos | (fneg << 16), fvo);
            for (int k0 = 0; k0 < n; k0 += 4) {
                const uint32_t x = *p++;
                for (int k = k0; k < n && k < k0 + 4; ++k) fv_ref[fvo + k] = o.x + ((x >> (8 * (k - k0))) & 255);
            }
            fvo += n;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K7 (MI): unique vertices + coordinates: compute_MI_vert_xyz (src/extract_mesh.cpp:1541-1637),
// barycentric coordinates from differences of adjacent material pairs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) write_verts_mi_kernel(const uint4* __restrict__ cand_key,
    const uint4* __restrict__ cand_pay, const uint32_t* __restrict__ rep, const uint32_t* __restrict__ vid,
    uint32_t n, const uint4* __restrict__ tets, const double* __restrict__ vals, uint32_t V,
    const double* __restrict__ pts, uint32_t* __restrict__ v_tet, uint8_t* __restrict__ v_local,
    uint8_t* __restrict__ v_size, uint4* __restrict__ v_simplex, uint4* __restrict__ v_funcs,
    double* __restrict__ v_xyz, uint4* __restrict__ v_key)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        if (rep[c] != c) continue;
        const uint32_t id = vid[c];
        const uint4 pay = cand_pay[c];
        v_key[id] = cand_key[c];
        const int size = (pay.y >> 8) & 255;
        uint32_t sv[4];
        if (size == 4) {
            const uint4 tv = __ldg(&tets[pay.x]);
            sv[0] = tv.x;
            sv[1] = tv.y;
            sv[2] = tv.z;
            sv[3] = tv.w;
        } else {
            const uint4 k = cand_key[c];
            sv[0] = k.x;
            sv[1] = k.y;
            sv[2] = k.z;
            sv[3] = NONE32;
        }
        uint32_t mi[4] = {pay.z & 0xffffu, pay.z >> 16, pay.w & 0xffffu, pay.w >> 16};
        for (int q = 0; q < 4; ++q)
            if (mi[q] == 0xffffu) mi[q] = NONE32;
        v_tet[id] = pay.x;
        v_local[id] = (uint8_t)(pay.y & 255);
        v_size[id] = (uint8_t)size;
        v_simplex[id] = make_uint4(sv[0], sv[1], sv[2], sv[3]);
        v_funcs[id] = make_uint4(mi[0], mi[1], mi[2], mi[3]);
        double out[3];
#define PT(v, c) pts[3 * (size_t)(v) + (c)]
#define FV(v, f) vals[(size_t)(f) * V + (v)]
        if (size == 1) {
            for (int d = 0; d < 3; ++d) out[d] = PT(sv[0], d);
        } else if (size == 2) {
            const double f1 = FV(sv[0], mi[0]) - FV(sv[0], mi[1]);
            const double f2 = FV(sv[1], mi[0]) - FV(sv[1], mi[1]);
            const double b0 = f2 / (f2 - f1), b1 = 1 - b0;
            for (int d = 0; d < 3; ++d) out[d] = b0 * PT(sv[0], d) + b1 * PT(sv[1], d);
        } else if (size == 3) {
            double p1[3], p2[3];
            for (int k = 0; k < 3; ++k) {
                p1[k] = FV(sv[k], mi[0]) - FV(sv[k], mi[1]);
                p2[k] = FV(sv[k], mi[1]) - FV(sv[k], mi[2]);
            }
            const double n1 = p1[2] * p2[1] - p1[1] * p2[2];
            const double n2 = p1[0] * p2[2] - p1[2] * p2[0];
            const double n3 = p1[1] * p2[0] - p1[0] * p2[1];
            const double dd = n1 + n2 + n3;
            const double w0 = n1 / dd, w1 = n2 / dd, w2 = n3 / dd;
            for (int d = 0; d < 3; ++d) out[d] = w0 * PT(sv[0], d) + w1 * PT(sv[1], d) + w2 * PT(sv[2], d);
        } else {
            double p1[4], p2[4], p3[4];
            for (int k = 0; k < 4; ++k) {
                p1[k] = FV(sv[k], mi[0]) - FV(sv[k], mi[1]);
                p2[k] = FV(sv[k], mi[1]) - FV(sv[k], mi[2]);
                p3[k] = FV(sv[k], mi[2]) - FV(sv[k], mi[3]);
            }
            const double n1 = p1[3] * (p2[2] * p3[1] - p2[1] * p3[2]) + p1[2] * (p2[1] * p3[3] - p2[3] * p3[1]) +
                              p1[1] * (p2[3] * p3[2] - p2[2] * p3[3]);
            const double n2 = p1[3] * (p2[0] * p3[2] - p2[2] * p3[0]) + p1[2] * (p2[3] * p3[0] - p2[0] * p3[3]) +
                              p1[0] * (p2[2] * p3[3] - p2[3] * p3[2]);
            const double n3 = p1[3] * (p2[1] * p3[0] - p2[0] * p3[1]) + p1[1] * (p2[0] * p3[3] - p2[3] * p3[0]) +
                              p1[0] * (p2[3] * p3[1] - p2[1] * p3[3]);
            const double n4 = p1[2] * (p2[0] * p3[1] - p2[1] * p3[0]) + p1[1] * (p2[2] * p3[0] - p2[0] * p3[2]) +
                              p1[0] * (p2[1] * p3[2] - p2[2] * p3[1]);
            const double dd = n1 + n2 + n3 + n4;
            const double w0 = n1 / dd, w1 = n2 / dd, w2 = n3 / dd, w3 = n4 / dd;
            for (int d = 0; d < 3; ++d)
                out[d] = w0 * PT(sv[0], d) + w1 * PT(sv[1], d) + w2 * PT(sv[2], d) + w3 * PT(sv[3], d);
        }
#undef PT
#undef FV
        v_xyz[3 * (size_t)id + 0] = out[0];
        v_xyz[3 * (size_t)id + 1] = out[1];
        v_xyz[3 * (size_t)id + 2] = out[2];
    }
}

// MI faces carry both labels: func_index = (positive material, negative material) (:831-832)
__global__ void __launch_bounds__(256) write_faces_mi_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
    uint32_t n_fv, uint32_t* __restrict__ f_off, uint32_t* __restrict__ f_toff, uint32_t* __restrict__ f_tets,
    uint32_t* __restrict__ f_funcs)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        if (i == n) {
            f_off[n] = n_fv;
            f_toff[n] = n;
            continue;
        }
        const uint4 h = face_hdr[i];
        f_off[i] = h.w;
        f_toff[i] = i;
        f_tets[2 * (size_t)i] = h.x;
        f_tets[2 * (size_t)i + 1] = h.y & 0xffffu;
        f_funcs[2 * (size_t)i] = h.z & 0xffffu;
        f_funcs[2 * (size_t)i + 1] = h.z >> 16;
    }
}

} // namespace rin
