// N1: edges of the extracted mesh on the device (compute_mesh_edges, src/mesh_connectivity.cpp:10-56).
// An edge is the unordered vertex pair of two consecutive entries of a face loop; its id is the order of its first
// occurrence scanning (face, position in face), i.e. the position p in the face-vertex array - the same
// hash-min + rank scheme as the vertex deduplication, with key (vmin, vmax) and candidate index p.
// Outputs: edge_verts[e] = (v1 <= v2), edges_of_face[p] = edge of entry p, and per edge the (face, position) pairs
// in ascending order (Edge::face_edge_indices, src/mesh.h:62-71).
#pragma once
#include "common.cuh"

namespace rin {

__global__ void __launch_bounds__(256) edge_keys_kernel(const uint32_t* __restrict__ f_off,
    const uint32_t* __restrict__ f_verts, uint32_t n_faces, uint2* __restrict__ ekey)
{
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n_faces; f += gridDim.x * blockDim.x) {
        const uint32_t b = f_off[f], e = f_off[f + 1];
        uint32_t prev = f_verts[b];
        const uint32_t first = prev;
        for (uint32_t p = b; p < e; ++p) {
            const uint32_t next = (p + 1 == e) ? first : f_verts[p + 1];
            ekey[p] = make_uint2(min(prev, next), max(prev, next));
            prev = next;
        }
    }
}

__device__ __forceinline__ uint32_t hash2(uint2 k)
{
    uint32_t h = k.x * 0x9E3779B1u;
    h = (h ^ (h >> 15)) + k.y * 0x85EBCA77u;
    h ^= h >> 13;
    h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) edge_insert_kernel(const uint2* __restrict__ ekey, uint32_t n,
    uint32_t* __restrict__ table, uint32_t mask, uint32_t* __restrict__ slot_of)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint2 k = ekey[p];
        uint32_t h = hash2(k) & mask;
        for (;;) {
            const uint32_t cur = atomicCAS(&table[h], NONE32, p);
            if (cur == NONE32) break;
            const uint2 kc = ekey[cur];
            if (kc.x == k.x && kc.y == k.y) {
                atomicMin(&table[h], p);
                break;
            }
            h = (h + 1) & mask;
        }
        slot_of[p] = h;
    }
}

// representatives (first occurrences) ranked in entry order; the slot then holds VID_FLAG | edge id
__global__ void __launch_bounds__(256) edge_rank_kernel(const uint2* __restrict__ ekey, uint32_t n,
    uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of, uint32_t* __restrict__ edge_verts,
    volatile unsigned long long* __restrict__ status, unsigned* __restrict__ tile_counter, unsigned* __restrict__ n_edges)
{
    __shared__ unsigned s_tile, s_base;
    __shared__ unsigned s_warp[8];
    constexpr int ITEMS = 4;
    const uint32_t n_tiles = (n + 256 * ITEMS - 1) / (256 * ITEMS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= n_tiles) return;
        const uint32_t base = tile * 256 * ITEMS + threadIdx.x * ITEMS;
        bool rep[ITEMS];
        uint32_t slot[ITEMS];
        unsigned cnt = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t p = base + j;
            rep[j] = false;
            slot[j] = 0;
            if (p < n) {
                slot[j] = slot_of[p];
                rep[j] = __ldcg(&table[slot[j]]) == p;
                cnt += rep[j];
            }
        }
        unsigned x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const unsigned t = (lane < 8) ? s_warp[lane] : 0;
            unsigned x8 = t;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, x8, o);
                if (lane >= o) x8 += y;
            }
            if (lane < 8) s_warp[lane] = x8 - t;
            const unsigned run = __shfl_sync(0xffffffffu, x8, 7);
            uint32_t e0, e1;
            tile_lookback_warp(status, (int)tile, run, 0, e0, e1);
            if (lane == 0) {
                s_base = e0;
                if (tile == n_tiles - 1) *n_edges = e0 + run;
            }
        }
        __syncthreads();
        unsigned id = s_base + s_warp[warp] + x - cnt;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
            if (rep[j]) {
                const uint2 k = ekey[base + j];
                edge_verts[2 * (size_t)id] = k.x;
                edge_verts[2 * (size_t)id + 1] = k.y;
                table[slot[j]] = VID_FLAG | id;
                ++id;
            }
    }
}

// edges_of_face + incidence counts
__global__ void __launch_bounds__(256) edge_assign_kernel(uint32_t n, const uint32_t* __restrict__ table,
    const uint32_t* __restrict__ slot_of, uint32_t* __restrict__ edges_of_face, uint32_t* __restrict__ cnt)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t e = table[slot_of[p]] & ~VID_FLAG;
        edges_of_face[p] = e;
        atomicAdd(&cnt[e], 1u);
    }
}

// exclusive scan of a uint32 array (tile look-back); out has n + 1 entries
__global__ void __launch_bounds__(256) scan_u32_kernel(const uint32_t* __restrict__ in, uint32_t n,
    uint32_t* __restrict__ out, volatile unsigned long long* __restrict__ status, unsigned* __restrict__ tile_counter)
{
    __shared__ unsigned s_tile, s_base;
    __shared__ unsigned s_warp[8];
    constexpr int ITEMS = 4;
    const uint32_t n_tiles = (n + 256 * ITEMS - 1) / (256 * ITEMS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= n_tiles) return;
        const uint32_t base = tile * 256 * ITEMS + threadIdx.x * ITEMS;
        uint32_t v[ITEMS];
        unsigned cnt = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            v[j] = (base + j < n) ? in[base + j] : 0u;
            cnt += v[j];
        }
        unsigned x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const unsigned t = (lane < 8) ? s_warp[lane] : 0;
            unsigned x8 = t;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, x8, o);
                if (lane >= o) x8 += y;
            }
            if (lane < 8) s_warp[lane] = x8 - t;
            const unsigned run = __shfl_sync(0xffffffffu, x8, 7);
            uint32_t e0, e1;
            tile_lookback_warp(status, (int)tile, run, 0, e0, e1);
            if (lane == 0) {
                s_base = e0;
                if (tile == n_tiles - 1) out[n] = e0 + run;
            }
        }
        __syncthreads();
        unsigned run = s_base + s_warp[warp] + x - cnt;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
            if (base + j < n) {
                out[base + j] = run;
                run += v[j];
            }
    }
}

// (face, position) pairs of every edge: appended in arbitrary order, then sorted per edge (lists are short)
__global__ void __launch_bounds__(256) edge_fill_kernel(const uint32_t* __restrict__ f_off, uint32_t n_faces,
    const uint32_t* __restrict__ edges_of_face, const uint32_t* __restrict__ eoff, uint32_t* __restrict__ cursor,
    uint2* __restrict__ pairs)
{
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n_faces; f += gridDim.x * blockDim.x) {
        const uint32_t b = f_off[f], e = f_off[f + 1];
        for (uint32_t p = b; p < e; ++p) {
            const uint32_t ed = edges_of_face[p];
            pairs[eoff[ed] + atomicAdd(&cursor[ed], 1u)] = make_uint2(f, p - b);
        }
    }
}

__global__ void __launch_bounds__(256) edge_sort_kernel(uint32_t n_edges, const uint32_t* __restrict__ eoff,
    uint2* __restrict__ pairs)
{
    for (uint32_t ed = blockIdx.x * blockDim.x + threadIdx.x; ed < n_edges; ed += gridDim.x * blockDim.x) {
        const uint32_t b = eoff[ed], e = eoff[ed + 1];
        for (uint32_t x = b + 1; x < e; ++x) {
            const uint2 v = pairs[x];
            uint32_t y = x;
            while (y > b && (pairs[y - 1].x > v.x || (pairs[y - 1].x == v.x && pairs[y - 1].y > v.y))) {
                pairs[y] = pairs[y - 1];
                --y;
            }
            pairs[y] = v;
        }
    }
}

} // namespace rin
