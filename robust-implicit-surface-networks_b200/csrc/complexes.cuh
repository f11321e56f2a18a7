// rin_get_complexes: full per-tet complexes on demand.
// The host topology stages of the reference read cut_results[cut_result_index[tet]] for a handful
// of tets (/root/reference/src/pair_faces.cpp:138-238, src/topo_ray_shooting.cpp:56-57,529-707);
// instead of materialising every complex, the big-tier kernels are re-run for the requested tets
// and the complete structure (Arrangement<3> / MaterialInterface<3> fields) is serialised.
#pragma once
#include "kernels_mi.cuh"

#include <type_traits>

namespace rin {

struct ComplexCounters
{
    unsigned top;      // words used in the output arena
    unsigned overflow; // an output did not fit
    int err;
    unsigned err_tet;
};

// index of tet t in the (ascending) active list, or NONE32
__device__ __forceinline__ uint32_t find_active(const uint32_t* __restrict__ act_tet, uint32_t n, uint32_t t)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (act_tet[mid] < t)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < n && act_tet[lo] == t) ? lo : NONE32;
}

template <int W>
__global__ void __launch_bounds__(GEN_THREADS) complexes_ia_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap, uint32_t n_active,
    const double* __restrict__ vals, uint32_t V, const uint32_t* __restrict__ req, uint32_t n_req,
    uint32_t* __restrict__ out, uint32_t out_cap, uint2* __restrict__ span, ComplexCounters* __restrict__ cc)
{
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_req; g += gridDim.x * blockDim.x) {
        const uint32_t t = req[g];
        const uint32_t a = find_active(act_tet, n_active, t);
        if (a == NONE32) {
            span[g] = make_uint2(0, 0); // cut_result_index == None
            continue;
        }
        const uint4 tv = __ldg(&tets[t]);
        IAComplex<IACaps> cx;
        cx.init();
        for (int w = 0; w < W; ++w) {
            uint32_t mm = act_mask[(size_t)w * cap + a];
            while (mm) {
                int f = w * 32 + __ffs(mm) - 1;
                mm &= mm - 1;
                double pv[4] = {__ldg(&vals[(size_t)f * V + tv.x]), __ldg(&vals[(size_t)f * V + tv.y]),
                    __ldg(&vals[(size_t)f * V + tv.z]), __ldg(&vals[(size_t)f * V + tv.w])};
                cx.insert(pv);
            }
        }
        if (cx.err) {
            if (atomicCAS(&cc->err, 0, cx.err == 1 ? RIN_ERR_CAPACITY : RIN_ERR_ARRANGEMENT) == 0) cc->err_tet = t;
            span[g] = make_uint2(0, 0);
            continue;
        }
        uint32_t words = 4 + 3 * cx.nv;
        for (int f = 0; f < cx.nf; ++f) words += 4 + cx.flen[f];
        for (int c = 0; c < cx.nc; ++c) words += 1 + cx.clen[c];
        if (cx.has_coplanar) words += 1 + 2 * cx.np;
        const uint32_t off = atomicAdd(&cc->top, words);
        span[g] = make_uint2(off, words);
        if (off + words > out_cap) {
            cc->overflow = 1;
            continue;
        }
        uint32_t* w = out + off;
        *w++ = cx.nv;
        *w++ = cx.nf;
        *w++ = cx.nc;
        *w++ = cx.has_coplanar ? cx.n_groups : 0;
        for (int v = 0; v < cx.nv; ++v) {
            *w++ = cx.vp[v][0];
            *w++ = cx.vp[v][1];
            *w++ = cx.vp[v][2];
        }
        for (int f = 0; f < cx.nf; ++f) {
            *w++ = cx.fplane[f];
            *w++ = cx.fpos[f] == N8 ? NONE32 : cx.fpos[f];
            *w++ = cx.fneg[f] == N8 ? NONE32 : cx.fneg[f];
            *w++ = cx.flen[f];
            for (int k = 0; k < cx.flen[f]; ++k) *w++ = cx.fv[cx.foff[f] + k];
        }
        for (int c = 0; c < cx.nc; ++c) {
            *w++ = cx.clen[c];
            for (int k = 0; k < cx.clen[c]; ++k) *w++ = cx.cf[cx.coff[c] + k];
        }
        if (cx.has_coplanar) {
            *w++ = cx.np;
            for (int p = 0; p < cx.np; ++p) *w++ = cx.upi[p];
            for (int p = 0; p < cx.np; ++p) *w++ = cx.same_orientation(p) ? 1u : 0u;
        }
    }
}

template <int W>
__global__ void __launch_bounds__(GEN_THREADS) complexes_mi_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap, uint32_t n_active,
    const double* __restrict__ vals, uint32_t V, const uint32_t* __restrict__ req, uint32_t n_req,
    uint32_t* __restrict__ out, uint32_t out_cap, uint2* __restrict__ span, ComplexCounters* __restrict__ cc)
{
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_req; g += gridDim.x * blockDim.x) {
        const uint32_t t = req[g];
        const uint32_t a = find_active(act_tet, n_active, t);
        if (a == NONE32) {
            span[g] = make_uint2(0, 0);
            continue;
        }
        const uint4 tv = __ldg(&tets[t]);
        MIComplex<MICaps> cx;
        bool first = true;
        for (int w = 0; w < W; ++w) {
            uint32_t mm = act_mask[(size_t)w * cap + a];
            while (mm) {
                int f = w * 32 + __ffs(mm) - 1;
                mm &= mm - 1;
                double pv[4] = {__ldg(&vals[(size_t)f * V + tv.x]), __ldg(&vals[(size_t)f * V + tv.y]),
                    __ldg(&vals[(size_t)f * V + tv.z]), __ldg(&vals[(size_t)f * V + tv.w])};
                if (first) {
                    cx.init(pv);
                    first = false;
                } else
                    cx.insert(pv);
            }
        }
        if (cx.err) {
            if (atomicCAS(&cc->err, 0, cx.err == 1 ? RIN_ERR_CAPACITY : RIN_ERR_ARRANGEMENT) == 0) cc->err_tet = t;
            span[g] = make_uint2(0, 0);
            continue;
        }
        const int B = cx.cur;
        uint32_t words = 4 + 4 * cx.nv;
        for (int f = 0; f < cx.nf; ++f) words += 3 + cx.flen[B][f];
        for (int c = 0; c < cx.nc; ++c) {
            words += 2;
            for (int f = 0; f < cx.nf; ++f) words += (cx.fpos[B][f] == c) + (cx.fneg[B][f] == c);
        }
        if (cx.has_dup) words += 1 + cx.nm;
        const uint32_t off = atomicAdd(&cc->top, words);
        span[g] = make_uint2(off, words);
        if (off + words > out_cap) {
            cc->overflow = 1;
            continue;
        }
        uint32_t* w = out + off;
        *w++ = cx.nv;
        *w++ = cx.nf;
        *w++ = cx.nc;
        *w++ = cx.has_dup ? cx.n_groups : 0;
        for (int v = 0; v < cx.nv; ++v)
            for (int q = 0; q < 4; ++q) *w++ = cx.vm[v][q];
        for (int f = 0; f < cx.nf; ++f) {
            *w++ = cx.pos_label(f);
            *w++ = cx.neg_label(f);
            *w++ = cx.flen[B][f];
            for (int k = 0; k < cx.flen[B][f]; ++k) *w++ = cx.fv[B][cx.foff[B][f] + k];
        }
        for (int c = 0; c < cx.nc; ++c) {
            *w++ = cx.cmat[c];
            uint32_t* cnt = w++;
            uint32_t n = 0;
            for (int f = 0; f < cx.nf; ++f) {
                if (cx.fpos[B][f] == c) {
                    *w++ = f;
                    ++n;
                }
                if (cx.fneg[B][f] == c) {
                    *w++ = f;
                    ++n;
                }
            }
            *cnt = n;
        }
        if (cx.has_dup) {
            *w++ = cx.nm;
            for (int p = 0; p < cx.nm; ++p) *w++ = cx.umi[p];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// robust_test (-R) of the reference (src/implicit_arrangement.cpp:137-243,
// src/material_interface.cpp:167-288): every active tet is computed twice, with the functions in
// forward and in reversed order; a failure of the first run is "type 2", of the second "type 3",
// different vertex / face / cell counts "type 1".
// ---------------------------------------------------------------------------------------------
struct RobustCounters
{
    unsigned type1, type2, type3, tested;
};

template <int W, bool MI>
__global__ void __launch_bounds__(GEN_THREADS) robust_test_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap, uint32_t n_active,
    const double* __restrict__ vals, uint32_t V, RobustCounters* __restrict__ rc)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const uint4 tv = __ldg(&tets[act_tet[a]]);
        uint32_t m[W];
        int k = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            m[w] = act_mask[(size_t)w * cap + a];
            k += __popc(m[w]);
        }
        int cnt[2][3];
        int err[2];
        for (int pass = 0; pass < 2; ++pass) {
            typename std::conditional<MI, MIComplex<MICaps>, IAComplex<IACaps>>::type cx;
            bool first = true;
            if (!MI) reinterpret_cast<IAComplex<IACaps>*>(&cx)->init();
            for (int j = 0; j < k; ++j) {
                const int f = nth_set_bit(m, W, pass ? k - 1 - j : j);
                double pv[4] = {__ldg(&vals[(size_t)f * V + tv.x]), __ldg(&vals[(size_t)f * V + tv.y]),
                    __ldg(&vals[(size_t)f * V + tv.z]), __ldg(&vals[(size_t)f * V + tv.w])};
                if (MI) {
                    MIComplex<MICaps>* c = reinterpret_cast<MIComplex<MICaps>*>(&cx);
                    if (first)
                        c->init(pv);
                    else
                        c->insert(pv);
                    first = false;
                } else
                    reinterpret_cast<IAComplex<IACaps>*>(&cx)->insert(pv);
            }
            err[pass] = cx.err;
            cnt[pass][0] = cx.nv;
            cnt[pass][1] = cx.nf;
            cnt[pass][2] = cx.nc;
        }
        atomicAdd(&rc->tested, 1u);
        if (err[0])
            atomicAdd(&rc->type2, 1u);
        else if (err[1])
            atomicAdd(&rc->type3, 1u);
        else if (cnt[0][0] != cnt[1][0] || cnt[0][1] != cnt[1][1] || cnt[0][2] != cnt[1][2])
            atomicAdd(&rc->type1, 1u);
    }
}

} // namespace rin
