// sm_100a kernels of the material-interface hot path (reference: src/material_interface.cpp:53-447,
// src/extract_mesh.cpp:569-986, :1541-1637).  Shares the scan / dedup / face kernels with the
// implicit-arrangement path (kernels_ia.cuh).
#pragma once
#include "kernels_ia.cuh"
#include "mi_complex.cuh"
#include "mi_complex_warp.cuh"

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
    uint2* __restrict__ tile_cnt, FilterCounters* __restrict__ ctr)
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
    unsigned k1 = 0, k2 = 0, km = 0, kf = 0;
#pragma unroll
    for (int j = 0; j < FILT_ITEMS; ++j) {
        const uint32_t i = base + j * FILT_THREADS + threadIdx.x;
        const uint4 tv = __ldg(&tets[t_first + min(i, t_count - 1)]);
        const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
        int kq = 0, ns = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            uint32_t s = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint2 g = __ldg(&vmask[(size_t)w * V + vv[c]]);
                s |= g.x;

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
    }
    if (lane == 0) s_kf[warp] = kf;
    __syncthreads();
    if (lane == 0) {
        if (k1) atomicAdd(&s_k[0], k1);
        if (k2) atomicAdd(&s_k[1], k2);
        if (km) atomicAdd(&s_k[2], km);
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
    int n_bf, n_extra, bf_words, bf_fv; // gated records: simplex-boundary faces
    __device__ static bool is_corner(const MIComplex<Caps>& cx, int v) { return cx.vm[v][2] < 4; }
    // number of materials identical to material m inside this tet (unique_materials group)
    __device__ static int group_size(const MIComplex<Caps>& cx, int m)
    {
        if (!cx.has_dup) return 1;
        int g = 0;
        for (int p = 4; p < cx.nm; ++p) g += (cx.umi[p] == cx.umi[m]);
        return g;
    }
    __device__ bool is_mi_vert(int v) const { return (isov[v >> 5] >> (v & 31)) & 1; }
    // boundary faces (positive label <= 3) of a tet whose materials tie on a whole tet face:
    // the extraction matches them with the neighbouring tet (src/extract_mesh.cpp:833-981)
    __device__ void run_boundary(const MIComplex<Caps>& cx)
    {
        const int B = cx.cur;
        n_bf = n_extra = bf_words = bf_fv = 0;
        for (int f = 0; f < cx.nf; ++f)
            if (!cx.is_mi_face(f)) {
                const int n = cx.flen[B][f];
                ++n_bf;
                bf_words += 2 + n;
                bf_fv += n;
                const int g = group_size(cx, cx.neg_label(f));
                if (g > 1) bf_words += (g + 3) / 4; // the inside cell carries several identical materials
                for (int k = 0; k < n; ++k) {
                    const int v = cx.fv[B][cx.foff[B][f] + k];
                    if (!is_mi_vert(v)) ++n_extra; // a tet corner not on the interface
                }
            }
    }
    __device__ void run(const MIComplex<Caps>& cx)
    {
        const int B = cx.cur;
        n_bf = n_extra = bf_words = bf_fv = 0;
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
    __device__ uint32_t size_bytes(bool gated) const
    {
        // non-gated records end with the face count of the whole complex (gated ones list every face)
        return 4u * uint32_t(1 + 2 * nvi + nfw + (gated ? 1 + bf_words : 1));
    }
    __device__ void write(const MIComplex<Caps>& cx, uint32_t* w, bool gated) const
    {
        const int B = cx.cur;
        int p = 0;
        if (gated) {
            w[p++] = (uint32_t)(nvi + n_extra) | ((uint32_t)(nfi + n_bf) << 8) | ((uint32_t)(nfv + bf_fv) << 16);
            w[p++] = (uint32_t)nvi | ((uint32_t)nfi << 8) | ((uint32_t)n_extra << 16);
        } else
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
        if (!gated) {
            w[p++] = (uint32_t)cx.nf; // trailing word (read by the cell-grouping maps only)
            return;
        }
        for (int f = 0; f < cx.nf; ++f)
            if (!cx.is_mi_face(f)) {
                const int n = cx.flen[B][f];
                const int inside = cx.neg_label(f);
                const int g = group_size(cx, inside);
                w[p++] = (uint32_t)f | ((uint32_t)n << 24);
                w[p++] = (uint32_t)cx.pos_label(f) | ((uint32_t)inside << 8) | ((uint32_t)g << 16);
                for (int k = 0; k < n; ++k) {
                    const int v = cx.fv[B][cx.foff[B][f] + k];
                    uint32_t x = ((uint32_t)v << 8);
                    if (is_corner(cx, v)) {
                        // the corner not among the three boundary pseudo materials
                        const int corner = 6 - cx.vm[v][0] - cx.vm[v][1] - cx.vm[v][2];
                        x |= 0x80000000u | ((uint32_t)corner << 16);
                    }
                    if (is_mi_vert(v))
                        x |= (uint32_t)rank(v);
                    else
                        x |= 0x40000000u;
                    w[p++] = x;
                }
                if (g > 1) { // members of the inside cell's material group, four ids per word
                    uint32_t x = 0;
                    int cnt = 0;
                    for (int q = 4; q < cx.nm; ++q) {
                        if (cx.umi[q] != cx.umi[inside]) continue;
                        x |= (uint32_t)q << (8 * (cnt & 3));
                        if ((++cnt & 3) == 0) {
                            w[p++] = x;
                            x = 0;
                        }
                    }
                    if (cnt & 3) w[p++] = x;
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
    const bool gated = (rec_ref[a] & REF_GATED) != 0;
    MIIsoScan<Caps> iso;
    if (!cx.err) {
        iso.run(cx);
        if (gated) {
            iso.run_boundary(cx);
        }
        if (iso.nvi + iso.n_extra > 255 || iso.nfi + iso.n_bf > 255 || iso.nfv + iso.bf_fv > 65535) cx.err = 1;
    }
    if (cx.n_exact) atomicAdd(&gc->n_exact, cx.n_exact);
    if (cx.err == 1 && !last_tier) return false;
    if (cx.err) {
        if (atomicCAS(&gc->err, 0, cx.err == 1 ? RIN_ERR_CAPACITY : RIN_ERR_ARRANGEMENT) == 0)
            gc->err_tet = act_tet[a];
        rec_ref[a] = REF_GENERAL;
        return true;
    }
    const uint32_t szal = iso.size_bytes(gated);
    const uint32_t off = atomicAdd(&gc->arena_top, szal);
    if (off + szal > arena_cap) {
        gc->arena_overflow = 1;
        rec_ref[a] = REF_GENERAL | (gated ? REF_GATED : 0u); // keep the flag for the retry
        return true;
    }
    iso.write(cx, reinterpret_cast<uint32_t*>(arena + off), gated);
    rec_ref[a] = REF_GENERAL | (gated ? REF_GATED : 0u) | (off >> 2);
    return true;
}

// Warp-cooperative version (small tier): the insertions run on all 32 lanes (mi_complex_warp.cuh), lane 0
// serialises the record.  All lanes call it with the same arguments.
template <class Caps, int W>
__device__ bool general_mi_one_warp(MIComplex<Caps>& cx, MIWarpScratch<Caps>& sc, uint32_t a,
    const uint4* __restrict__ tets, const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask,
    uint32_t cap, const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc, bool last_tier, int lane, int skip = 0)
{
    const uint4 tv = __ldg(&tets[act_tet[a]]);
    bool first = skip == 0; // skip > 0: the complex already holds the first `skip` materials (tabulated start)
    for (int w = 0; w < W; ++w) {
        uint32_t mm = act_mask[(size_t)w * cap + a];
        while (mm) {
            const int f = w * 32 + __ffs(mm) - 1;
            mm &= mm - 1;
            if (skip > 0) {
                --skip;
                continue;
            }
            double pv[4];
            pv[0] = __ldg(&vals[(size_t)f * V + tv.x]);
            pv[1] = __ldg(&vals[(size_t)f * V + tv.y]);
            pv[2] = __ldg(&vals[(size_t)f * V + tv.z]);
            pv[3] = __ldg(&vals[(size_t)f * V + tv.w]);
            if (first) {
                if (lane == 0) cx.init(pv);
                __syncwarp();
                first = false;
            } else
                warp_insert_material(cx, sc, pv, lane);
        }
    }
    __syncwarp();
    int done = 1;
    if (lane == 0) {
        const bool gated = (rec_ref[a] & REF_GATED) != 0;
        MIIsoScan<Caps> iso;
        if (!cx.err) {
            iso.run(cx);
            if (gated) iso.run_boundary(cx);
            if (iso.nvi + iso.n_extra > 255 || iso.nfi + iso.n_bf > 255 || iso.nfv + iso.bf_fv > 65535) cx.err = 1;
        }
        if (cx.n_exact) atomicAdd(&gc->n_exact, cx.n_exact);
        if (cx.err == 1 && !last_tier)
            done = 0;
        else if (cx.err) {
            if (atomicCAS(&gc->err, 0, cx.err == 1 ? RIN_ERR_CAPACITY : RIN_ERR_ARRANGEMENT) == 0)
                gc->err_tet = act_tet[a];
            rec_ref[a] = REF_GENERAL;
        } else {
            const uint32_t szal = iso.size_bytes(gated);
            const uint32_t off = atomicAdd(&gc->arena_top, szal);
            if (off + szal > arena_cap) {
                gc->arena_overflow = 1;
                rec_ref[a] = REF_GENERAL | (gated ? REF_GATED : 0u); // keep the flag for the retry
            } else {
                iso.write(cx, reinterpret_cast<uint32_t*>(arena + off), gated);
                rec_ref[a] = REF_GENERAL | (gated ? REF_GATED : 0u) | (off >> 2);
            }
        }
    }
    done = __shfl_sync(0xffffffffu, done, 0);
    __syncwarp();
    return done != 0;
}

__device__ __forceinline__ int mi3_key(const double p0[4], const double p1[4], const double p2[4], unsigned* nex);

struct alignas(16) MISmallSlot
{
    MIComplex<MICapsSmall> cx;
    MIWarpScratch<MICapsSmall> sc;
};

template <int W>
__global__ void __launch_bounds__(GEN_SMALL_WARPS * 32) general_mi_small_kernel(
    const uint4* __restrict__ tets, const uint32_t* __restrict__ act_tet,
    const uint32_t* __restrict__ act_mask, uint32_t cap, const uint32_t* __restrict__ small_list,
    uint32_t* __restrict__ ovf_list, const double* __restrict__ vals, uint32_t V,
    uint8_t* __restrict__ arena, uint32_t arena_cap, uint32_t* __restrict__ rec_ref,
    GeneralCounters* __restrict__ gc, const MIComplex<MICapsSmall>* __restrict__ cx3 = nullptr,
    const uint32_t* __restrict__ lut3cx = nullptr)
{
    extern __shared__ __align__(16) uint8_t s_raw_mi[];
    MISmallSlot* s_slot = reinterpret_cast<MISmallSlot*>(s_raw_mi);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = gc->n_small;
    for (uint32_t g = blockIdx.x * GEN_SMALL_WARPS + warp; g < n; g += gridDim.x * GEN_SMALL_WARPS) {
        const uint32_t a = small_list[g];
        int skip = 0;
        if (cx3) {
            // Tabulated start (the analogue of the 2-plane start of general_ia_small_kernel): a tet with >= 4 materials
            // whose first three have a tabulated key starts from the complete complex of those three and inserts only
            // the rest.  Every branch of an insertion depends on exact signs alone, which the key determines
            // (mi3_key), so the state equals what the insertions would have produced; the values are the tet's own.
            int entry = -1;
            double p0[4], p1[4], p2[4];
            if (lane == 0 && !(rec_ref[a] & REF_GATED)) {
                uint32_t m[W];
                int kk = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    m[w] = act_mask[(size_t)w * cap + a];
                    kk += __popc(m[w]);
                }
                if (kk >= 4) {
                    const uint4 tv = __ldg(&tets[act_tet[a]]);
                    const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
                    const int f0 = nth_set_bit(m, W, 0), f1 = nth_set_bit(m, W, 1), f2 = nth_set_bit(m, W, 2);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        p0[c] = __ldg(&vals[(size_t)f0 * V + vv[c]]);
                        p1[c] = __ldg(&vals[(size_t)f1 * V + vv[c]]);
                        p2[c] = __ldg(&vals[(size_t)f2 * V + vv[c]]);
                    }
                    unsigned ex = 0;
                    const int key = mi3_key(p0, p1, p2, &ex);
                    if (ex) atomicAdd(&gc->n_exact, ex);
                    if (key >= 0 && lut3cx[key] != 0xffffffffu) entry = (int)lut3cx[key];
                }
            }
            entry = __shfl_sync(0xffffffffu, entry, 0);
            if (entry >= 0) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(cx3 + entry);
                uint32_t* dst = reinterpret_cast<uint32_t*>(&s_slot[warp].cx);
                for (int i = lane; i < (int)(sizeof(MIComplex<MICapsSmall>) / 4); i += 32) dst[i] = src[i];
                __syncwarp();
                if (lane == 0) {
                    MIComplex<MICapsSmall>& cx = s_slot[warp].cx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        cx.mval[0][c] = p0[c];
                        cx.mval[1][c] = p1[c];
                        cx.mval[2][c] = p2[c];
                    }
                    cx.n_exact = 0;
                    cx.err = 0;
                }
                skip = 3;
            }
        }
        __syncwarp();
        if (!general_mi_one_warp<MICapsSmall, W>(s_slot[warp].cx, s_slot[warp].sc, a, tets, act_tet, act_mask, cap,
                vals, V, arena, arena_cap, rec_ref, gc, false, lane, skip)) {
            if (lane == 0) ovf_list[atomicAdd(&gc->n_ovf, 1u)] = a;
        }
        __syncwarp();
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
// Key of the interface of THREE materials (the reference's "secondary" table, dispatch at
// src/material_interface.cpp:320-324).  The general algorithm starts with m0, inserts m1, then m2; every branch
// it takes depends on exact signs only: the order of the three values at each corner (3 bits per corner) and,
// for every tet edge on which m0 - m1 changes sign, the sign of m2 - m0 at the point of that edge where
// m0 = m1 (6 bits; it is the vertex the first insertion created there).  Equal keys -> identical complexes.
// -1: a tie at a corner or a vanishing edge predicate (degenerate: general kernel).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int mi3_key(const double p0[4], const double p1[4], const double p2[4], unsigned* nex)
{
    int key = 0;
    int a01 = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (p0[c] == p1[c] || p0[c] == p2[c] || p1[c] == p2[c]) return -1;
        const int a = p0[c] > p1[c], b = p0[c] > p2[c], d = p1[c] > p2[c];
        key |= (a | (b << 1) | (d << 2)) << (3 * c);
        a01 |= a << c;
    }
    int e = 0;
    for (int x = 0; x < 4; ++x)
        for (int y = x + 1; y < 4; ++y, ++e) {
            if (((a01 >> x) & 1) == ((a01 >> y) & 1)) continue;
            // exact sign of (m2 - m0) where m0 = m1 on edge (x, y): the predicate of MIComplex::orient_vertex
            // for the vertex {boundary, boundary, m0, m1}
            const double qa[4] = {p0[x], p0[y], p2[x], p2[y]}, qb[4] = {p1[x], p1[y], p0[x], p0[y]};
            const double da[4] = {p0[x], p0[y], 1.0, 1.0}, db[4] = {p1[x], p1[y], 0.0, 0.0};
            const int sq = detn_diff_sign(2, qa, qb, nex);
            if (sq == 0) return -1;
            const int sd = detn_diff_sign(2, da, db, nex);
            if (sd == 0) return -1;
            if (sq * sd > 0) key |= 1 << (12 + e);
        }
    return key;
}
constexpr uint32_t MI3_KEYS = 1u << 18;
constexpr uint32_t LUT3_MISS = 0xffffffffu;

// dumps the complete 3-material complexes of the chosen witnesses (table generation; witness w owns vertices
// 4w..4w+3 and materials 0..2): the start state of general_mi_small_kernel for tets with more materials
__global__ void __launch_bounds__(GEN_THREADS) dump_mi3_kernel(const uint32_t* __restrict__ witness, uint32_t n,
    const double* __restrict__ vals, uint32_t V, MIComplex<MICapsSmall>* __restrict__ out, int* __restrict__ err)
{
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        const uint32_t w = witness[g];
        MIComplex<MICapsSmall> cx;
        for (int f = 0; f < 3; ++f) {
            double pv[4];
            for (int c = 0; c < 4; ++c) pv[c] = vals[(size_t)f * V + 4 * w + c];
            if (f == 0)
                cx.init(pv);
            else
                cx.insert(pv);
        }
        if (cx.err) *err = cx.err;
        cx.n_exact = 0;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&cx);
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + g);
        for (int i = 0; i < (int)(sizeof(MIComplex<MICapsSmall>) / 4); ++i) dst[i] = src[i];
    }
}

// table generation: witness w owns vertices 4w..4w+3 and materials 0..2 with hashed values in (0, 1)
__global__ void __launch_bounds__(256) mi3_witness_kernel(uint32_t n, uint32_t Vw, double* __restrict__ vals,
    int* __restrict__ keys)
{
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
        double p[3][4];
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                unsigned long long z = ((unsigned long long)w * 12 + f * 4 + c + 1) * 0x9e3779b97f4a7c15ull;
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
                z ^= z >> 31;
                p[f][c] = (double)((z >> 11) + 1) * (1.0 / 9007199254740994.0);
                vals[(size_t)f * Vw + 4 * w + c] = p[f][c];
            }
        unsigned nex = 0;
        keys[w] = mi3_key(p[0], p[1], p[2], &nex);
    }
}

// ---------------------------------------------------------------------------------------------
// K3 (MI): two materials without ties go to the 16-entry table (sign pattern of m0 - m1 at the
// corners), three materials without degeneracies to the secondary table (mi3_key) when use_secondary_lookup is
// on; everything else to the general kernels.  Dispatch of src/material_interface.cpp:320-328.
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) classify_mi_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    uint32_t n_active, const double* __restrict__ vals, const uint2* __restrict__ vmask, uint32_t V, uint32_t F,
    const uint16_t* __restrict__ lut2, const uint32_t* __restrict__ lut3, int use_lookup, int use_secondary,
    uint32_t* __restrict__ rec_ref, uint32_t* __restrict__ small_list, uint32_t* __restrict__ big_list,
    GeneralCounters* __restrict__ gc, unsigned* __restrict__ n_gated, const unsigned* __restrict__ n_dev = nullptr)
{
    if (n_dev) n_active = min(n_active, *n_dev); // launched for the capacity: the count is still on the device
    unsigned exact = 0;
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        uint32_t m[W];
        int k = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            m[w] = act_mask[(size_t)w * cap + a];
            k += __popc(m[w]);
        }
        uint32_t ref = REF_GENERAL;
        const uint4 tv = __ldg(&tets[act_tet[a]]);
        const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
        // tie gate: an active material and any other material coincide exactly at the three corners
        // of one tet face -> a simplex-boundary face piece may be a material interface
        bool gated = false;
        {
            unsigned tie_cnt = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) tie_cnt += __ldg(&vmask[vv[c]]).y & 1u;
            if (tie_cnt >= 3) {
                for (uint32_t g = 0; g < F && !gated; ++g) {
                    bool active = false;
#pragma unroll
                    for (int w = 0; w < W; ++w)
                        if ((g >> 5) == (uint32_t)w) active = (m[w] >> (g & 31)) & 1;
                    if (!active) continue;
                    double xg[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) xg[c] = __ldg(&vals[(size_t)g * V + vv[c]]);
                    for (uint32_t f = 0; f < F && !gated; ++f) {
                        if (f == g) continue;
                        int eq = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) eq += (__ldg(&vals[(size_t)f * V + vv[c]]) == xg[c]);
                        gated = (eq >= 3);
                    }
                }
            }
        }
        if (gated) {
            ref = REF_GENERAL | REF_GATED;
            atomicAdd(n_gated, 1u);
        } else if (use_lookup && k == 2) {
            const int f0 = nth_set_bit(m, W, 0), f1 = nth_set_bit(m, W, 1);
            int key = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double x = __ldg(&vals[(size_t)f0 * V + vv[c]]), y = __ldg(&vals[(size_t)f1 * V + vv[c]]);
                if (x == y) key = -1000;
                key |= (x > y ? 1 : 0) << c;
            }
            if (key >= 0) ref = lut2[key];
        } else if (use_lookup && use_secondary && k == 3 && lut3) {
            // the 3-material ("secondary") table, src/material_interface.cpp:320-324
            const int f0 = nth_set_bit(m, W, 0), f1 = nth_set_bit(m, W, 1), f2 = nth_set_bit(m, W, 2);
            double p0[4], p1[4], p2[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                p0[c] = __ldg(&vals[(size_t)f0 * V + vv[c]]);
                p1[c] = __ldg(&vals[(size_t)f1 * V + vv[c]]);
                p2[c] = __ldg(&vals[(size_t)f2 * V + vv[c]]);
            }
            const int key = mi3_key(p0, p1, p2, &exact);
            if (key >= 0) {
                const uint32_t off = __ldg(&lut3[key]);
                if (off != LUT3_MISS) ref = off;
            }
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
    if (exact) atomicAdd(&gc->n_exact, exact);
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
    uint4* __restrict__ cand_pay, uint4* __restrict__ face_hdr, uint32_t* __restrict__ fv_ref,
    uint32_t* __restrict__ bf_mask /* [face slot][W] material sets of boundary faces; null without tie tets */,
    unsigned* __restrict__ n_bface_slots)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const uint32_t ref = rec_ref[a];
        const uint32_t* r = (ref & REF_GENERAL)
                                ? reinterpret_cast<const uint32_t*>(arena) + (size_t)(ref & ~REF_FLAGS)
                                : reinterpret_cast<const uint32_t*>(lut_blob) + ref;
        const bool gated = (ref & REF_GATED) && (ref & ~REF_FLAGS); // offset 0 = empty error record
        uint32_t hdr = r[0];
        if ((hdr & 0xffffu) == 0) continue;
        int n_total_f = (hdr >> 8) & 255;
        if (gated) hdr = r[1]; // (MI vertices, interface faces, reserved corner slots)
        const int nv = hdr & 255, nf = (hdr >> 8) & 255;
        const uint32_t t = act_tet[a];
        const uint4 tv4 = __ldg(&tets[t]);
        const uint32_t tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
        uint32_t m[W];
#pragma unroll
        for (int w = 0; w < W; ++w) m[w] = act_mask[(size_t)w * cap + a];
        const uint4 o = offs[a];
        const uint32_t* p = r + (gated ? 2 : 1);
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
            // gated records list EVERY face of the complex: the slot is the local face id, so that the
            // output keeps the reference's local face order (interface and boundary faces interleave)
            face_hdr[o.y + (gated ? local : (uint32_t)j)] =
                make_uint4(t, local | ((uint32_t)n << 16), fpos | (fneg << 16), fvo);
            for (int k0 = 0; k0 < n; k0 += 4) {
                const uint32_t x = *p++;
                for (int k = k0; k < n && k < k0 + 4; ++k) fv_ref[fvo + k] = o.x + ((x >> (8 * (k - k0))) & 255);
            }
            fvo += n;
        }
        if (!gated) continue;
        // simplex-boundary faces of a tie tet: reserved (inactive) face slots and, for tet corners
        // that are not interface vertices, reserved (dead) vertex candidates; both are switched on
        // by mi_bface_decide_kernel when the neighbouring tet holds a different material
        uint32_t xe = o.x + nv;
        const int nbf = n_total_f - nf;
        for (int j = 0; j < nbf; ++j) {
            const uint32_t e0 = *p++, e1 = *p++;
            const uint32_t local = e0 & 0xffffu;
            const int n = (e0 >> 24) & 127;
            const uint32_t bface = e1 & 255;
            const uint32_t inside = (uint32_t)nth_set_bit(m, W, (int)((e1 >> 8) & 255) - 4);
            const int grp = (e1 >> 16) & 255;
            face_hdr[o.y + local] =
                make_uint4(t, local | ((uint32_t)n << 16) | FACE_BND | FACE_INACTIVE, bface | (inside << 16), fvo);
            for (int k = 0; k < n; ++k) {
                const uint32_t x = *p++;
                if (x & 0x40000000u) { // corner that is not an interface vertex
                    const uint32_t corner = (x >> 16) & 3;
                    cand_key[xe] = make_uint4(tv[corner], NONE32, NONE32, 0x3fffffffu);
                    cand_pay[xe] = make_uint4(t, ((x >> 8) & 255) | (1u << 8) | CAND_DEAD, 0xffffffffu, 0xffffffffu);
                    fv_ref[fvo + k] = xe++;
                } else
                    fv_ref[fvo + k] = o.x + (x & 255);
            }
            // the set of materials the inside cell stands for (several when materials coincide)
            uint32_t lm[W];
#pragma unroll
            for (int w = 0; w < W; ++w) lm[w] = 0;
            if (grp > 1) {
                for (int q0 = 0; q0 < grp; q0 += 4) {
                    const uint32_t x = *p++;
                    for (int q = q0; q < grp && q < q0 + 4; ++q) {
                        const uint32_t gm = (uint32_t)nth_set_bit(m, W, (int)((x >> (8 * (q - q0))) & 255) - 4);
#pragma unroll
                        for (int w = 0; w < W; ++w)
                            if ((gm >> 5) == (uint32_t)w) lm[w] |= 1u << (gm & 31);
                    }
                }
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w)
                    if ((inside >> 5) == (uint32_t)w) lm[w] |= 1u << (inside & 31);
            }
            if (bf_mask)
#pragma unroll
                for (int w = 0; w < W; ++w) bf_mask[(size_t)(o.y + local) * W + w] = lm[w];
            fvo += n;
        }
        if (nbf) atomicAdd(n_bface_slots, (unsigned)nbf);
    }
}

// ---------------------------------------------------------------------------------------------
// Degenerate ties only: matching of simplex-boundary faces between neighbouring tets
// (src/extract_mesh.cpp:833-981).  A boundary face is identified by (smallest, second smallest,
// largest) vertex identity, where tet corners count as corners (the reference's -(id)-1) and
// every other vertex by its first-occurrence candidate.  The SECOND tet (in tet order) that sees a
// face whose inside material differs from the first tet's emits it as a material-interface face.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mi_bface_keys_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
    const uint32_t* __restrict__ fv_ref, const uint4* __restrict__ cand_key, const uint4* __restrict__ cand_pay,
    const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ table, uint4* __restrict__ fkeys)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = face_hdr[i];
        if (!(h.y & FACE_BND)) continue;
        const int nv = (h.y >> 16) & 255;
        uint32_t mn = 0, mx = 0, second = 0;
        int mn_pos = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (int k = 0; k < nv; ++k) {
                const uint32_t c = fv_ref[h.w + k];
                uint32_t id;
                if (((cand_pay[c].y >> 8) & 255) == 1)
                    id = 0x80000000u | cand_key[c].x; // on a tet vertex
                else {
                    const uint32_t s = slot_of[c];
                    id = (s == NONE32) ? c : table[s];
                }
                if (pass == 0) {
                    if (k == 0) {
                        mn = mx = id;
                        mn_pos = 0;
                    } else if (id < mn) {
                        mn = id;
                        mn_pos = k;
                    } else if (id > mx)
                        mx = id;
                } else {
                    if (k == 0) second = 0xffffffffu;
                    if (k != mn_pos && id < second) second = id;
                }
            }
        fkeys[i] = make_uint4(mn, second, mx, 0);
    }
}

__global__ void __launch_bounds__(256) mi_bface_decide_kernel(uint4* __restrict__ face_hdr, uint32_t n,
    const uint32_t* __restrict__ frep, const uint32_t* __restrict__ ndup, const uint32_t* __restrict__ fv_ref,
    uint4* __restrict__ cand_pay, const uint32_t* __restrict__ bf_mask, const uint32_t* __restrict__ act_tet,
    const uint32_t* __restrict__ act_mask, uint32_t cap, uint32_t n_active, int W, unsigned* __restrict__ n_bad,
    uint32_t* __restrict__ partner /* first visitor's slot -> the slot that carries the face */)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint4 h = face_hdr[i];
        if (!(h.y & FACE_BND)) continue;
        const uint32_t r = frep[i];
        if (r == i) continue; // first visitor (or unmatched): stays inactive
        if (ndup[r] > 1) {
            atomicAdd(n_bad, 1u); // more than two tets on one face: not a manifold tet mesh
            continue;
        }
        const uint32_t mine = h.z >> 16;
        uint32_t common = 0; // a material present on both sides: not an interface (:870-950)
        for (int w = 0; w < W; ++w) common |= bf_mask[(size_t)i * W + w] & bf_mask[(size_t)r * W + w];
        if (common) continue;
        // func_index.first = material_in_tet[positive label - 4 + start] with a boundary label (< 4):
        // QUIRK kept from the reference (:979): an earlier CRS entry
        uint32_t first = 0xffffu;
        {
            uint32_t lo = 0, hi = n_active;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (act_tet[mid] < h.x)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            int back = 4 - (int)(h.z & 255);
            for (uint32_t ap = lo; ap > 0 && back > 0;) {
                --ap;
                int kp = 0;
                for (int w = 0; w < W; ++w) kp += __popc(act_mask[(size_t)w * cap + ap]);
                if (back <= kp) {
                    int want = kp - back; // index among the set bits
                    for (int w = 0; w < W; ++w) {
                        const uint32_t mm = act_mask[(size_t)w * cap + ap];
                        const int c = __popc(mm);
                        if (want < c) {
                            first = w * 32 + __fns(mm, 0, want + 1);
                            break;
                        }
                        want -= c;
                    }
                    back = 0;
                } else
                    back -= kp;
            }
        }
        h.y &= ~FACE_INACTIVE;
        h.z = first | (mine << 16);
        face_hdr[i] = h;
        partner[r] = i;
        const int nv = (h.y >> 16) & 255;
        for (int k = 0; k < nv; ++k) {
            const uint32_t c = fv_ref[h.w + k];
            uint4 p = cand_pay[c];
            if (p.y & CAND_DEAD) {
                p.y = (p.y & ~CAND_DEAD) | CAND_DEDUP;
                cand_pay[c] = p;
            }
        }
    }
}

// keep[i] as a representative array for bface_scan / bface_write: i when the face stays, NONE32 else
__global__ void __launch_bounds__(256) mi_face_keep_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
    uint32_t* __restrict__ keep)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        keep[i] = (face_hdr[i].y & FACE_INACTIVE) ? NONE32 : i;
}

// ---------------------------------------------------------------------------------------------
// Cell-grouping maps, material-interface version (second extract_MI_mesh overload,
// src/extract_mesh.cpp:988-1443): global_vId_of_tet_vert as for the arrangement (corners -(id)-1,
// :1130,1259) and MI_fId_of_tet_face; a simplex-boundary face that became an interface between two tie
// tets carries the same face id in BOTH tets (:1391-1392).
// ---------------------------------------------------------------------------------------------
struct MIRecView
{
    const uint32_t* verts; // nvi entries of 2 words
    const uint32_t* faces; // interface faces
    int nvi, nfi, n_bf, nf_total;
    bool gated;
};
__device__ __forceinline__ MIRecView mi_rec_view(uint32_t ref, const uint8_t* lut_blob, const uint8_t* arena)
{
    const uint32_t* r = (ref & REF_GENERAL) ? reinterpret_cast<const uint32_t*>(arena) + (size_t)(ref & ~REF_FLAGS)
                                            : reinterpret_cast<const uint32_t*>(lut_blob) + ref;
    MIRecView v;
    v.gated = (ref & REF_GATED) && (ref & ~REF_FLAGS);
    const uint32_t hdr = v.gated ? r[1] : r[0];
    v.nvi = hdr & 255;
    v.nfi = (hdr >> 8) & 255;
    v.verts = r + (v.gated ? 2 : 1);
    v.faces = v.verts + 2 * v.nvi;
    v.n_bf = v.gated ? (int)((r[0] >> 8) & 255) - v.nfi : 0;
    v.nf_total = 0;
    return v;
}
// end of the interface-face entries
__device__ __forceinline__ const uint32_t* mi_rec_skip_faces(const MIRecView& v)
{
    const uint32_t* p = v.faces;
    for (int q = 0; q < v.nfi; ++q) p += 1 + rec_face_words((p[0] >> 24) & 127);
    return p;
}

__global__ void __launch_bounds__(256) tetmap_mi_count_kernel(const uint32_t* __restrict__ rec_ref,
    uint32_t n_active, const uint8_t* __restrict__ lut_blob, const uint8_t* __restrict__ arena,
    uint2* __restrict__ cnt)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const MIRecView v = mi_rec_view(rec_ref[a], lut_blob, arena);
        int top = 4;
        for (int i = 0; i < v.nvi; ++i) top = max(top, (int)(v.verts[2 * i] & 255) + 1);
        const uint32_t* p = mi_rec_skip_faces(v);
        cnt[a] = make_uint2((uint32_t)top, v.gated ? (uint32_t)(v.nfi + v.n_bf) : p[0]);
    }
}

// degenerate runs (frep != nullptr): frep[slot] = slot when the face was kept, fpos = its final position,
// partner[slot] = the slot of the neighbouring tet that carries the shared boundary face
__global__ void __launch_bounds__(256) tetmap_mi_write_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, uint32_t n_active, const uint32_t* __restrict__ rec_ref,
    const uint4* __restrict__ offs, const uint8_t* __restrict__ lut_blob, const uint8_t* __restrict__ arena,
    const uint32_t* __restrict__ rep, const uint32_t* __restrict__ vid, const uint32_t* __restrict__ frep,
    const uint4* __restrict__ fpos, const uint32_t* __restrict__ partner, const uint2* __restrict__ off,
    long long* __restrict__ vmap, uint32_t* __restrict__ fmap)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const MIRecView v = mi_rec_view(rec_ref[a], lut_blob, arena);
        const uint4 tv4 = __ldg(&tets[act_tet[a]]);
        const uint32_t tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
        const uint4 o = offs[a];
        const uint2 b = off[a], e = off[a + 1];
        for (int j = 0; j < 4; ++j) vmap[b.x + j] = -(long long)tv[j] - 1;
        for (int i = 0; i < v.nvi; ++i) {
            const int local = v.verts[2 * i] & 255;
            if (local >= 4) vmap[b.x + local] = (long long)vid[rep[o.x + i]];
        }
        for (uint32_t f = b.y; f < e.y; ++f) fmap[f] = NONE32;
        auto final_id = [&](uint32_t slot) -> uint32_t {
            if (!frep) return slot;
            if (frep[slot] == slot) return fpos[slot].x;
            const uint32_t q = partner[slot];
            return (q != NONE32 && frep[q] == q) ? fpos[q].x : NONE32;
        };
        const uint32_t* p = v.faces;
        for (int q = 0; q < v.nfi; ++q) {
            const uint32_t local = p[0] & 0xffffu;
            fmap[b.y + local] = final_id(o.y + (v.gated ? local : (uint32_t)q));
            p += 1 + rec_face_words((p[0] >> 24) & 127);
        }
        for (int q = 0; q < v.n_bf; ++q) { // gated records: every simplex-boundary face has a slot
            const uint32_t local = p[0] & 0xffffu;
            const int n = (p[0] >> 24) & 127, g = (p[1] >> 16) & 255;
            fmap[b.y + local] = final_id(o.y + local);
            p += 2 + n + (g > 1 ? (g + 3) / 4 : 0);
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
