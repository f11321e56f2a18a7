// Iso record of a finished per-tet complex: the serial writer (one thread per tet, big tier) and the
// warp-cooperative writer (one tet per warp).  Both emit the byte-identical record (record_words.cuh).
#pragma once
#include "ia_complex_warp.cuh"
#include "record_words.cuh"

namespace rin {

// iso part of a finished complex: counts, then serialisation straight into the arena
template <class Caps>
struct IsoScan
{
    uint32_t isov[(Caps::MAXV + 31) / 32];
    int nvi, nfi, nfv, nfw; // nfw: words used by the face entries
    int nbnd;               // iso faces on the tet boundary (negative_cell == None)
    __device__ void run(const IAComplex<Caps>& cx)
    {
        for (int i = 0; i < (Caps::MAXV + 31) / 32; ++i) isov[i] = 0;
        nfi = 0;
        nfv = 0;
        nfw = 0;
        nbnd = 0;
        for (int f = 0; f < cx.nf; ++f)
            if (cx.is_iso_face(f)) {
                ++nfi;
                nbnd += (cx.fneg[f] == N8);
                nfv += cx.flen[f];
                if (cx.flen[f] > 127) nfv = 1 << 20; // loop too long for the record format -> capacity error
                nfw += rec_face_words(cx.flen[f]);
                for (int k = 0; k < cx.flen[f]; ++k) {
                    int v = cx.fv[cx.foff[f] + k];
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
    __device__ uint32_t size_bytes() const { return 4u * uint32_t(1 + nvi + nfw + 1); }
    __device__ void write(const IAComplex<Caps>& cx, uint32_t* w) const
    {
        w[1 + nvi + nfw] = (uint32_t)cx.nf; // trailing word: faces of the whole complex (record_words.cuh)
        int p = 0;
        w[p++] = (uint32_t)nvi | ((uint32_t)nfi << 8) | ((uint32_t)nfv << 16);
        for (int v = 0; v < cx.nv; ++v)
            if ((isov[v >> 5] >> (v & 31)) & 1)
                w[p++] = (uint32_t)v | ((uint32_t)cx.vp[v][0] << 8) | ((uint32_t)cx.vp[v][1] << 16) |
                         ((uint32_t)cx.vp[v][2] << 24);
        for (int f = 0; f < cx.nf; ++f)
            if (cx.is_iso_face(f)) {
                const int n = cx.flen[f];
                w[p++] = (uint32_t)f | ((uint32_t)cx.fplane[f] << 16) | ((uint32_t)n << 24) |
                         ((cx.fneg[f] == N8) ? 0x80000000u : 0u);
                for (int k0 = 0; k0 < n; k0 += 4) {
                    uint32_t x = 0;
                    for (int k = k0; k < n && k < k0 + 4; ++k)
                        x |= (uint32_t)rank(cx.fv[cx.foff[f] + k]) << (8 * (k - k0));
                    w[p++] = x;
                }
            }
    }
};

} // namespace rin

namespace rin {

// iso part of a finished complex, computed and serialised by the whole warp (same record as IsoScan)
template <class Caps>
struct WarpIso
{
    static constexpr int NW = (Caps::MAXV + 31) / 32;
    uint32_t isov[NW]; // identical in all lanes after run()
    int nvi, nfi, nfv, nfw, nbnd;

    __device__ void run(const IAComplex<Caps>& cx, int lane)
    {
#pragma unroll
        for (int i = 0; i < NW; ++i) isov[i] = 0;
        int lfi = 0, lfv = 0, lfw = 0, lbnd = 0;
        bool toolong = false;
        for (int f = lane; f < cx.nf; f += 32)
            if (cx.is_iso_face(f)) {
                const int n = cx.flen[f], off = cx.foff[f];
                ++lfi;
                lbnd += (cx.fneg[f] == N8);
                lfv += n;
                toolong |= n > 127;
                lfw += (int)rec_face_words(n);
                for (int k = 0; k < n; ++k) {
                    const int v = cx.fv[off + k];
                    isov[v >> 5] |= 1u << (v & 31);
                }
            }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lfi += __shfl_xor_sync(WFULL, lfi, d);
            lfv += __shfl_xor_sync(WFULL, lfv, d);
            lfw += __shfl_xor_sync(WFULL, lfw, d);
            lbnd += __shfl_xor_sync(WFULL, lbnd, d);
#pragma unroll
            for (int i = 0; i < NW; ++i) isov[i] |= __shfl_xor_sync(WFULL, isov[i], d);
        }
        nfi = lfi;
        nbnd = lbnd;
        nfv = __ballot_sync(WFULL, toolong) ? (1 << 20) : lfv; // loop too long for the record format
        nfw = lfw;
        nvi = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) nvi += __popc(isov[i]);
    }
    __device__ int rank(int v) const
    {
        int r = __popc(isov[v >> 5] & ((1u << (v & 31)) - 1u));
        for (int i = 0; i < (v >> 5); ++i) r += __popc(isov[i]);
        return r;
    }
    __device__ uint32_t size_bytes() const { return 4u * uint32_t(1 + nvi + nfw + 1); }
    __device__ void write(const IAComplex<Caps>& cx, uint32_t* w, int lane) const
    {
        if (lane == 0) {
            w[0] = (uint32_t)nvi | ((uint32_t)nfi << 8) | ((uint32_t)nfv << 16);
            w[1 + nvi + nfw] = (uint32_t)cx.nf; // trailing word: faces of the whole complex
        }
        for (int v = lane; v < cx.nv; v += 32)
            if ((isov[v >> 5] >> (v & 31)) & 1)
                w[1 + rank(v)] = (uint32_t)v | ((uint32_t)cx.vp[v][0] << 8) | ((uint32_t)cx.vp[v][1] << 16) |
                                 ((uint32_t)cx.vp[v][2] << 24);
        int p = 1 + nvi;
        for (int base = 0; base < cx.nf; base += 32) {
            const int f = base + lane;
            const bool iso = (f < cx.nf) && cx.is_iso_face(f);
            const int n = iso ? (int)cx.flen[f] : 0;
            int tot;
            int o = p + warp_excl_scan(iso ? (int)rec_face_words(n) : 0, lane, tot);
            if (iso) {
                const int off = cx.foff[f];
                w[o++] = (uint32_t)f | ((uint32_t)cx.fplane[f] << 16) | ((uint32_t)n << 24) |
                         ((cx.fneg[f] == N8) ? 0x80000000u : 0u);
                for (int k0 = 0; k0 < n; k0 += 4) {
                    uint32_t x = 0;
                    for (int k = k0; k < n && k < k0 + 4; ++k)
                        x |= (uint32_t)rank(cx.fv[off + k]) << (8 * (k - k0));
                    w[o++] = x;
                }
            }
            p += tot;
        }
    }
};

} // namespace rin
