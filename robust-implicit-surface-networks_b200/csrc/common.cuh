// Shared device helpers: decoupled look-back tile scan, hashing, record format.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "record_words.cuh"

namespace rin {

constexpr uint32_t NONE32 = 0xffffffffu;

// per-tile totals / exclusive prefixes of the streaming filter (32 bytes): active tets, active functions
// (CRS length), vertex candidates, faces, face-vertex entries
struct TileTot
{
    unsigned act, funcs, cand, face, fv, pad0, pad1, pad2;
};

// ---------------------------------------------------------------------------------------------
// Decoupled look-back over tiles for a PAIR of 31-bit partial sums packed with a 2-bit flag in
// one 64-bit word, so that flag and value are published atomically.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long ST_AGG = 1ull << 62, ST_PRE = 2ull << 62;

__device__ __forceinline__ unsigned long long st_pack(unsigned long long flag, uint32_t a, uint32_t b)
{
    return flag | ((unsigned long long)(a & 0x7fffffffu) << 31) | (unsigned long long)(b & 0x7fffffffu);
}

// Called by ONE thread of the tile.  Publishes this tile's aggregate, sums the predecessors and
// publishes the inclusive prefix.  Returns the exclusive prefix in (ea, eb).
__device__ __forceinline__ void tile_lookback(
    volatile unsigned long long* st, int tile, uint32_t a, uint32_t b, uint32_t& ea, uint32_t& eb)
{
    ea = 0;
    eb = 0;
    if (tile > 0) {
        st[tile] = st_pack(ST_AGG, a, b);
        for (int p = tile - 1; p >= 0; --p) {
            unsigned long long w;
            do {
                w = st[p];
            } while ((w >> 62) == 0);
            ea += uint32_t((w >> 31) & 0x7fffffffu);
            eb += uint32_t(w & 0x7fffffffu);
            if ((w >> 62) == 2) break;
        }
    }
    st[tile] = st_pack(ST_PRE, ea + a, eb + b);
}

// Warp-parallel variant: called by all 32 lanes of ONE warp with the same (a, b); each step
// inspects 32 predecessors at once.  Returns the exclusive prefix in every lane.
__device__ __forceinline__ void tile_lookback_warp(
    volatile unsigned long long* st, int tile, uint32_t a, uint32_t b, uint32_t& ea, uint32_t& eb)
{
    const int lane = threadIdx.x & 31;
    uint32_t sa = 0, sb = 0;
    if (tile > 0) {
        if (lane == 0) st[tile] = st_pack(ST_AGG, a, b);
        int p = tile - 1;
        for (;;) {
            const int idx = p - lane;
            unsigned long long w = ST_PRE; // before the first tile: prefix 0
            if (idx >= 0) {
                do {
                    w = st[idx];
                } while ((w >> 62) == 0);
            }
            const unsigned pre = __ballot_sync(0xffffffffu, (w >> 62) == 2);
            const int upto = pre ? (__ffs(pre) - 1) : 31; // nearest predecessor holding a prefix
            uint32_t va = (lane <= upto) ? uint32_t((w >> 31) & 0x7fffffffu) : 0u;
            uint32_t vb = (lane <= upto) ? uint32_t(w & 0x7fffffffu) : 0u;
            for (int o = 16; o; o >>= 1) {
                va += __shfl_xor_sync(0xffffffffu, va, o);
                vb += __shfl_xor_sync(0xffffffffu, vb, o);
            }
            sa += va;
            sb += vb;
            if (pre) break;
            p -= 32;
        }
    }
    ea = sa;
    eb = sb;
    if (lane == 0) st[tile] = st_pack(ST_PRE, ea + a, eb + b);
}

__device__ __forceinline__ uint32_t hash4(uint4 k)
{
    uint32_t h = k.x * 0x9E3779B1u;
    h = (h ^ (h >> 15)) + k.y * 0x85EBCA77u;
    h = (h ^ (h >> 13)) + k.z * 0xC2B2AE3Du;
    h = (h ^ (h >> 16)) + k.w * 0x27D4EB2Fu;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}

// Division by a run-time constant without the XU pipe (I2F / MUFU.RCP / F2I): multiply-high by a magic
// number (the branch-free scheme of Granlund & Montgomery as popularised by libdivide).
struct FastDiv
{
    uint32_t m, s, d;
};
inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    f.d = d;
    if (d <= 1) {
        f.m = 0;
        f.s = 0xffffffffu;
        return f;
    }
    const uint32_t l = 31u - (uint32_t)__builtin_clz(d);
    if ((d & (d - 1)) == 0) {
        f.m = 0;
        f.s = l - 1;
        return f;
    }
    const uint64_t num = 1ull << (32 + l);
    uint64_t pm = num / d;
    const uint64_t rem = num % d;
    pm += pm;
    if (rem + rem >= d) pm += 1;
    f.m = (uint32_t)(pm + 1);
    f.s = l;
    return f;
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv f)
{
    if (f.s == 0xffffffffu) return n;
    const uint32_t q = __umulhi(n, f.m);
    return (((n - q) >> 1) + q) >> f.s;
}

// After the ranking pass of the implicit-arrangement pipeline a candidate's slot_of entry, or the table entry
// of its slot, holds VID_FLAG | final vertex id.
constexpr uint32_t VID_FLAG = 0x80000000u;
__device__ __forceinline__ uint32_t final_vid(uint32_t c, const uint32_t* __restrict__ slot_of,
    const uint32_t* __restrict__ table)
{
    const uint32_t s = slot_of[c];
    return ((s & VID_FLAG) ? s : table[s]) & ~VID_FLAG;
}

__device__ __forceinline__ bool key_eq(uint4 a, uint4 b)
{
    return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}

} // namespace rin
