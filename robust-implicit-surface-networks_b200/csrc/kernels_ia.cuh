// sm_100a kernels of the implicit-arrangement hot path (K1..K7 of SURVEY.md section 2-D).
// Bounds: K1/K2 stream V*F doubles and T index quadruples (HBM); K3/K5/K6/K7 touch only the
// ~7 % active tets; K4 is latency / local-memory bound.  No tensor cores: nothing is a contraction.
#pragma once
#include "../../include/rin_b200.h"
#include "common.cuh"
#include "ia_complex.cuh"
#include "iso_record.cuh"

namespace rin {

// ---------------------------------------------------------------------------------------------
// K0: generate_tet_mesh on the device (/root/reference/src/io.cpp:95-152)
// ---------------------------------------------------------------------------------------------
__global__ void grid_points_kernel(uint32_t N, double3 bmin, double3 bmax, double* __restrict__ pts)
{
    uint64_t n = (uint64_t)N * N * N;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n;
         v += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t k = v % N, j = (v / N) % N, i = v / ((uint64_t)N * N);
        double d = double(N - 1);
        // t * (max - min) + min with t = i / (N-1)   (src/io.cpp:104-113)
        pts[3 * v + 0] = (double(i) / d) * (bmax.x - bmin.x) + bmin.x;
        pts[3 * v + 1] = (double(j) / d) * (bmax.y - bmin.y) + bmin.y;
        pts[3 * v + 2] = (double(k) / d) * (bmax.z - bmin.z) + bmin.z;
    }
}

__constant__ uint8_t c_grid_even[5][4] = {{4, 6, 1, 3}, {6, 3, 4, 7}, {1, 3, 0, 4}, {3, 1, 2, 6}, {4, 1, 6, 5}};
__constant__ uint8_t c_grid_odd[5][4] = {{7, 0, 2, 5}, {2, 3, 0, 7}, {5, 7, 0, 4}, {7, 2, 6, 5}, {0, 1, 2, 5}};

__global__ void grid_tets_kernel(uint32_t R, uint4* __restrict__ tets)
{
    const uint32_t N = R + 1;
    uint64_t n = (uint64_t)R * R * R * 5;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t cube = t / 5;
        uint32_t s = t % 5;
        uint32_t k = cube % R, j = (cube / R) % R, i = cube / ((uint64_t)R * R);
        const uint8_t* tab = ((i + j + k) & 1) ? c_grid_odd[s] : c_grid_even[s];
        uint32_t v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t q = tab[c];
            // corner q of the cube, numbered like v0..v7 at src/io.cpp:126-133
            uint32_t di = ((q & 3) == 1 || (q & 3) == 2), dj = ((q & 3) >= 2), dk = (q >> 2);
            v[c] = (i + di) * N * N + (j + dj) * N + (k + dk);
        }
        tets[t] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void narrow_tets_kernel(const uint64_t* __restrict__ in, uint64_t n, uint4* __restrict__ out)
{
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + 4 * t);
        ulonglong2 a = p[0], b = p[1];
        out[t] = make_uint4((uint32_t)a.x, (uint32_t)a.y, (uint32_t)b.x, (uint32_t)b.y);
    }
}

// min / max vertex id referenced by a tet range (slab sharding evaluates only that vertex range)
__global__ void vertex_range_kernel(const uint4* __restrict__ tets, uint32_t t_first, uint32_t n,
    uint32_t* __restrict__ minmax)
{
    uint32_t lo = 0xffffffffu, hi = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 t = tets[t_first + i];
        lo = min(min(lo, t.x), min(min(t.y, t.z), t.w));
        hi = max(max(hi, t.x), max(max(t.y, t.z), t.w));
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
}

// ---------------------------------------------------------------------------------------------
// K1: evaluate every function at every vertex (SoA double[F][V]) + per-vertex sign bit masks.
// Replaces load_functions (app/implicit_arrangement.cpp:57-64) and the "func signs" loop
// (src/implicit_arrangement.cpp:61-77).  vmask[w*V + v] = (P bits, N bits) of functions 32w..32w+31.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double eval_func(const rin_func_desc& f, double x, double y, double z)
{
    double v = 0.0;
    switch (f.type) {
    case RIN_FN_PLANE: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        v = (f.p[3] * dx + f.p[4] * dy) + f.p[5] * dz;
        break;
    }
    case RIN_FN_SPHERE: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        double s = (dx * dx + dy * dy) + dz * dz;
        v = (f.p[4] != 0.0) ? f.p[3] * f.p[3] - s : f.p[3] - sqrt(s);
        break;
    }
    case RIN_FN_CYLINDER: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        double t = (f.p[3] * dx + f.p[4] * dy) + f.p[5] * dz;
        double px = dx - t * f.p[3], py = dy - t * f.p[4], pz = dz - t * f.p[5];
        v = f.p[6] - sqrt((px * px + py * py) + pz * pz);
        break;
    }
    case RIN_FN_TORUS: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        double t = (f.p[3] * dx + f.p[4] * dy) + f.p[5] * dz;
        double px = dx - t * f.p[3], py = dy - t * f.p[4], pz = dz - t * f.p[5];
        double rho = sqrt((px * px + py * py) + pz * pz) - f.p[6];
        v = f.p[7] - sqrt(rho * rho + t * t);
        break;
    }
    default: break;
    }
    return f.flip ? -v : v;
}

__global__ void __launch_bounds__(256) eval_functions_kernel(const double* __restrict__ pts,
    uint32_t v_first, uint32_t v_count, uint32_t V, const rin_func_desc* __restrict__ funcs, uint32_t F,
    int negate, double* __restrict__ vals, uint2* __restrict__ vmask, uint32_t* __restrict__ vmask16,
    unsigned long long* __restrict__ n_zero)
{
    // vmask16 (nullable, F <= 16): P | N << 16 in ONE word per vertex: halves the sectors the filter's
    // four mask gathers touch
    extern __shared__ rin_func_desc s_funcs[];
    for (uint32_t i = threadIdx.x; i < F * (sizeof(rin_func_desc) / 8); i += blockDim.x)
        reinterpret_cast<double*>(s_funcs)[i] = reinterpret_cast<const double*>(funcs)[i];
    __syncthreads();
    unsigned zeros = 0;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < v_count;
         idx += gridDim.x * blockDim.x) {
        const uint32_t v = v_first + idx;
        const double x = pts[3 * (size_t)v], y = pts[3 * (size_t)v + 1], z = pts[3 * (size_t)v + 2];
        for (uint32_t w = 0; w * 32 < F; ++w) {
            uint32_t P = 0, Nn = 0;
            const uint32_t fe = min(F, w * 32 + 32);
            for (uint32_t f = w * 32; f < fe; ++f) {
                double val = eval_func(s_funcs[f], x, y, z);
                if (negate) val = val * -1; // csg(): funcVals * -1 (src/csg.cpp:37)
                vals[(size_t)f * V + v] = val;
                P |= (val > 0 ? 1u : 0u) << (f & 31);
                Nn |= (val < 0 ? 1u : 0u) << (f & 31);
            }
            if (vmask16)
                vmask16[v] = P | (Nn << 16);
            else
                vmask[(size_t)w * V + v] = make_uint2(P, Nn);
            zeros += (fe - w * 32) - __popc(P | Nn);
        }
    }
    // num_degenerate_vertex counts (vertex, function) pairs with value 0 (:69-73)
    for (int o = 16; o; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(n_zero, (unsigned long long)zeros);
}

// values supplied by the caller as the reference's row-major V x F matrix: transpose to SoA and
// build the sign masks (the "func signs" loop).
__global__ void __launch_bounds__(256) ingest_values_kernel(const double* __restrict__ rowmajor,
    uint32_t v_first, uint32_t v_count, uint32_t V, uint32_t F, int negate, double* __restrict__ vals,
    uint2* __restrict__ vmask, uint32_t* __restrict__ vmask16, unsigned long long* __restrict__ n_zero)
{
    unsigned zeros = 0;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < v_count;
         idx += gridDim.x * blockDim.x) {
        const uint32_t v = v_first + idx;
        for (uint32_t w = 0; w * 32 < F; ++w) {
            uint32_t P = 0, Nn = 0;
            const uint32_t fe = min(F, w * 32 + 32);
            for (uint32_t f = w * 32; f < fe; ++f) {
                double val = rowmajor[(size_t)v * F + f];
                if (negate) val = val * -1;
                vals[(size_t)f * V + v] = val;
                P |= (val > 0 ? 1u : 0u) << (f & 31);
                Nn |= (val < 0 ? 1u : 0u) << (f & 31);
            }
            if (vmask16)
                vmask16[v] = P | (Nn << 16);
            else
                vmask[(size_t)w * V + v] = make_uint2(P, Nn);
            zeros += (fe - w * 32) - __popc(P | Nn);
        }
    }
    for (int o = 16; o; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(n_zero, (unsigned long long)zeros);
}

// ---------------------------------------------------------------------------------------------
// K2a: per-tet active-function filter + ordered stream compaction.
// Replaces the "filter" loop (src/implicit_arrangement.cpp:86-116): function j is active in a tet
// iff it is neither positive at all four vertices nor negative at all four (pos<4 && neg<4, :106).
// Output: active tets in tet order with their W-word function masks.
// One streaming pass over the 16-byte index records writes tile-local compacted slots (no
// inter-block dependency, hence no spinning warps); a one-block scan of the tile totals and a
// gather over the ~7 % active tets produce the ordered list.
// ---------------------------------------------------------------------------------------------
constexpr int FILT_THREADS = 256;
constexpr int FILT_ITEMS = 8;
constexpr int FILT_BATCH = 4;
constexpr int FILT_TILE = FILT_THREADS * FILT_ITEMS;

struct FilterCounters
{
    unsigned tile_counter;
    unsigned n_active;
    unsigned n_k1, n_k2, n_kmore;
    unsigned n_funcs; // sum of popcounts = |func_in_tet|
};

// Pass 1 (streams the index records): per-tile compaction into tile-local slots
// tl_tet / tl_mask [tile * FILT_TILE + rank] and per-tile totals.  No inter-block dependency.
template <int W, bool PACK>
__global__ void __launch_bounds__(FILT_THREADS) filter_tiles_kernel(const uint4* __restrict__ tets,
    uint32_t t_first, uint32_t t_count, const uint2* __restrict__ vmask, const uint32_t* __restrict__ vmask16,
    uint32_t V, uint32_t last_mask,
    uint32_t* __restrict__ tl_tet, uint32_t* __restrict__ tl_mask, size_t tl_stride,
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
    // batches of FILT_BATCH items: all index loads of a batch are issued before the first
    // dependent mask gather, all gathers before the first use (memory-level parallelism)
#pragma unroll
    for (int j0 = 0; j0 < FILT_ITEMS; j0 += FILT_BATCH) {
        uint4 tv[FILT_BATCH];
#pragma unroll
        for (int j = 0; j < FILT_BATCH; ++j) {
            const uint32_t i = base + (j0 + j) * FILT_THREADS + threadIdx.x;
            tv[j] = __ldg(&tets[t_first + min(i, t_count - 1)]);
        }
        int kk[FILT_BATCH];
#pragma unroll
        for (int j = 0; j < FILT_BATCH; ++j) kk[j] = 0;
        if (PACK) {
            // one 32-bit word per vertex: the AND of the four words holds "positive everywhere" in
            // the low half and "negative everywhere" in the high half
            uint32_t g[FILT_BATCH][4];
#pragma unroll
            for (int j = 0; j < FILT_BATCH; ++j) {
                g[j][0] = __ldg(&vmask16[tv[j].x]);
                g[j][1] = __ldg(&vmask16[tv[j].y]);
                g[j][2] = __ldg(&vmask16[tv[j].z]);
                g[j][3] = __ldg(&vmask16[tv[j].w]);
            }
#pragma unroll
            for (int j = 0; j < FILT_BATCH; ++j) {
                const uint32_t x = g[j][0] & g[j][1] & g[j][2] & g[j][3];
                uint32_t mm = ~((x & 0xffffu) | (x >> 16)) & last_mask;
                const uint32_t i = base + (j0 + j) * FILT_THREADS + threadIdx.x;
                if (i >= t_count) mm = 0;
                m[j0 + j][0] = mm;
                kk[j] += __popc(mm);
            }
        } else {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            uint2 g[FILT_BATCH][4];
#pragma unroll
            for (int j = 0; j < FILT_BATCH; ++j) {
                g[j][0] = __ldg(&vmask[(size_t)w * V + tv[j].x]);
                g[j][1] = __ldg(&vmask[(size_t)w * V + tv[j].y]);
                g[j][2] = __ldg(&vmask[(size_t)w * V + tv[j].z]);
                g[j][3] = __ldg(&vmask[(size_t)w * V + tv[j].w]);
            }
#pragma unroll
            for (int j = 0; j < FILT_BATCH; ++j) {
                uint32_t mm = ~((g[j][0].x & g[j][1].x & g[j][2].x & g[j][3].x) |
                                (g[j][0].y & g[j][1].y & g[j][2].y & g[j][3].y));
                if (w == W - 1) mm &= last_mask;
                const uint32_t i = base + (j0 + j) * FILT_THREADS + threadIdx.x;
                if (i >= t_count) mm = 0;
                m[j0 + j][w] = mm;
                kk[j] += __popc(mm);
            }
        }
        }
#pragma unroll
        for (int j = 0; j < FILT_BATCH; ++j) {
            const int kq = kk[j];
            k1 += (kq == 1);
            k2 += (kq == 2);
            km += (kq > 2);
            kf += kq;
            ball[j0 + j] = __ballot_sync(0xffffffffu, kq > 0);
            if (lane == 0) s_cnt[j0 + j][warp] = __popc(ball[j0 + j]);
        }
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
    // every warp derives the exclusive prefix over the (item, warp) sequence it needs by itself
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
        // entry (j, warp) lives in lane (j * 8 + warp) % 32 of round (j * 8 + warp) / 32
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

// Pass 2 (one block): exclusive scan of the per-tile totals -> tile_off[tile] = (first active
// index, first CRS index); tile_off[n_tiles] = totals.
__global__ void __launch_bounds__(1024) scan_tiles_kernel(const uint2* __restrict__ tile_cnt, uint32_t n_tiles,
    uint2* __restrict__ tile_off, FilterCounters* __restrict__ ctr)
{
    __shared__ uint2 s_w[32];
    __shared__ uint2 s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_tiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint2 c = (i < n_tiles) ? tile_cnt[i] : make_uint2(0, 0);
        uint2 x = c;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y0 = __shfl_up_sync(0xffffffffu, x.x, o), y1 = __shfl_up_sync(0xffffffffu, x.y, o);
            if (lane >= o) {
                x.x += y0;
                x.y += y1;
            }
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint2 t = s_w[lane], y = t;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned z0 = __shfl_up_sync(0xffffffffu, y.x, o), z1 = __shfl_up_sync(0xffffffffu, y.y, o);
                if (lane >= o) {
                    y.x += z0;
                    y.y += z1;
                }
            }
            s_w[lane] = make_uint2(y.x - t.x, y.y - t.y);
        }
        __syncthreads();
        const uint2 r = s_run, wv = s_w[warp];
        if (i < n_tiles) tile_off[i] = make_uint2(r.x + wv.x + x.x - c.x, r.y + wv.y + x.y - c.y);
        __syncthreads();
        if (threadIdx.x == 1023) s_run = make_uint2(r.x + wv.x + x.x, r.y + wv.y + x.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_off[n_tiles] = s_run;
        ctr->n_active = s_run.x;
        ctr->n_funcs = s_run.y;
    }
}

// Pass 3: ordered gather of the tile-local slots into the compact active list (7 % of the tets)
template <int W>
__global__ void __launch_bounds__(256) compact_active_kernel(const uint32_t* __restrict__ tl_tet,
    const uint32_t* __restrict__ tl_mask, size_t tl_stride, const uint2* __restrict__ tile_off,
    uint32_t n_tiles, uint32_t n_active, uint32_t* __restrict__ act_tet, uint32_t* __restrict__ act_mask,
    uint32_t cap, uint32_t tile_slots = FILT_TILE, const unsigned* __restrict__ n_dev = nullptr)
{
    if (n_dev) n_active = min(n_active, *n_dev); // launched for the capacity: the count is still on the device
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        // tile = last one with tile_off.x <= a
        uint32_t lo = 0, hi = n_tiles;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&tile_off[mid].x) <= a)
                lo = mid;
            else
                hi = mid;
        }
        const size_t pos = (size_t)lo * tile_slots + (a - __ldg(&tile_off[lo].x));
        act_tet[a] = tl_tet[pos];
#pragma unroll
        for (int w = 0; w < W; ++w) act_mask[(size_t)w * cap + a] = tl_mask[(size_t)w * tl_stride + pos];
    }
}

// ---------------------------------------------------------------------------------------------
// K3: classify active tets: table hit (record offset in the LUT blob) or general work list.
// Dispatch rules of src/implicit_arrangement.cpp:276-284: tables serve 1 function, and 2
// functions when the secondary lookup is on; any zero value at a tet vertex, coincident crossing
// points or a missing table entry go to the general kernel.
// rec_ref[a]: bit31 = general (payload filled by K4), else 4-byte-unit offset into the LUT blob.
// ---------------------------------------------------------------------------------------------
struct GeneralCounters
{
    unsigned n_general; // all tets that take a general kernel
    unsigned n_small;   // ... of which <= IACapsSmall::MAXK functions (shared-memory tier)
    unsigned n_big;     // ... big tier (more functions)
    unsigned n_ovf;     // small-tier capacity overflows, re-queued for the big tier
    unsigned arena_top; // bytes
    unsigned n_exact;
    int err;            // first error code (RIN_ERR_*)
    unsigned err_tet;
    unsigned arena_overflow;
    unsigned n_ovf2;    // mid-tier capacity overflows (IA), re-queued for the per-thread big tier
    unsigned n_bnd_faces; // iso faces on a tet boundary among the general records (degenerate inputs)
    unsigned done_blocks; // last-block ticket of the mid tier (tile scan in its tail)
};

constexpr uint32_t REF_GENERAL = 0x80000000u;
constexpr uint32_t REF_GATED = 0x40000000u; // MI: record also lists the simplex-boundary faces (degenerate ties)
constexpr uint32_t REF_FLAGS = REF_GENERAL | REF_GATED;
constexpr uint16_t LUT_MISS = 0xffffu;

struct LutView
{
    const uint16_t* lut1; // [16]   record offset (4-byte units) per vertex-sign pattern
    const uint16_t* lut2; // [256*64] outer | inner<<8
    const uint8_t* blob;
    uint32_t blob_bytes;
};

__device__ __forceinline__ int nth_set_bit(const uint32_t* m, int W, int n)
{
    for (int w = 0; w < W; ++w) {
        int c = __popc(m[w]);
        if (n < c) return w * 32 + __fns(m[w], 0, n + 1);
        n -= c;
    }
    return -1;
}

// key of a 2-plane arrangement: vertex signs (8 bits) | order of the two crossing points on every
// simplex edge crossed by both planes (6 bits, << 8); -1 when a value is zero or two crossings coincide
__device__ __forceinline__ int ia2_key(const double p0[4], const double p1[4], unsigned* exact)
{
    int s0 = 0, s1 = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (p0[c] == 0.0 || p1[c] == 0.0) return -1;
        s0 |= (p0[c] > 0 ? 1 : 0) << c;
        s1 |= (p1[c] > 0 ? 1 : 0) << c;
    }
    int key = s0 | (s1 << 4);
    int e = 0;
    for (int x = 0; x < 4; ++x)
        for (int y = x + 1; y < 4; ++y, ++e) {
            bool c0 = ((s0 >> x) & 1) != ((s0 >> y) & 1);
            bool c1 = ((s1 >> x) & 1) != ((s1 >> y) & 1);
            if (!(c0 && c1)) continue;
            int s = det2_sign(p0[x], p0[y], p1[x], p1[y], exact);
            if (s == 0) return -1;
            if (s > 0) key |= 1 << (8 + e);
        }
    return key;
}

template <int W>
__global__ void __launch_bounds__(256) classify_ia_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    uint32_t n_active, const uint2* __restrict__ vmask, const double* __restrict__ vals, uint32_t V,
    LutView lut, int use_lookup, int use_secondary, uint32_t* __restrict__ rec_ref,
    uint32_t* __restrict__ small_list, uint32_t* __restrict__ big_list, GeneralCounters* __restrict__ gc,
    unsigned* __restrict__ n_exact, int* __restrict__ key_out)
{
    __shared__ uint16_t s_lut2[256 * 64];
    __shared__ uint16_t s_lut1[16];
    if (use_lookup) {
        for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) s_lut2[i] = lut.lut2[i];
        if (threadIdx.x < 16) s_lut1[threadIdx.x] = lut.lut1[threadIdx.x];
    }
    __syncthreads();
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
        int key = -1;
        if (use_lookup && (k == 1 || (k == 2 && (use_secondary || key_out)))) {
            const uint4 tv = __ldg(&tets[act_tet[a]]);
            const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
            const int f0 = nth_set_bit(m, W, 0);
            const int f1 = (k == 2) ? nth_set_bit(m, W, 1) : f0;
            int s0 = 0, s1 = 0;
            bool nonzero = true;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint2 a0 = __ldg(&vmask[(size_t)(f0 >> 5) * V + vv[c]]);
                s0 |= ((a0.x >> (f0 & 31)) & 1) << c;
                nonzero &= (((a0.x | a0.y) >> (f0 & 31)) & 1) != 0;
                if (k == 2) {
                    const uint2 a1 = __ldg(&vmask[(size_t)(f1 >> 5) * V + vv[c]]);
                    s1 |= ((a1.x >> (f1 & 31)) & 1) << c;
                    nonzero &= (((a1.x | a1.y) >> (f1 & 31)) & 1) != 0;
                }
            }
            if (nonzero) {
                if (k == 1) {
                    key = s0;
                    ref = s_lut1[s0];
                } else {
                    double p0[4], p1[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        p0[c] = __ldg(&vals[(size_t)f0 * V + vv[c]]);
                        p1[c] = __ldg(&vals[(size_t)f1 * V + vv[c]]);
                    }
                    key = ia2_key(p0, p1, &exact);
                    if (key >= 0 && use_secondary) {
                        uint16_t off = s_lut2[key];
                        ref = (off == LUT_MISS) ? REF_GENERAL : off;
                    }
                }
            }
        }
        if (key_out) key_out[a] = key;
        if (ref & REF_GENERAL) {
            atomicAdd(&gc->n_general, 1u);
            if (k <= IACapsSmall::MAXK)
                small_list[atomicAdd(&gc->n_small, 1u)] = a;
            else
                big_list[atomicAdd(&gc->n_big, 1u)] = a;
        }
        rec_ref[a] = ref;
    }
    if (exact) atomicAdd(n_exact, exact);
}

// ---------------------------------------------------------------------------------------------
// K4: general arrangement kernel: one thread per tet, complex in local memory, iso record out.
// ---------------------------------------------------------------------------------------------

// One general tet: gather the active functions' values, build the complex, publish the record.
// Returns false when the complex did not fit this tier (caller re-queues it for the big tier).
template <class Caps, int W>
__device__ bool general_ia_one(IAComplex<Caps>& cx, uint32_t a, const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc, bool last_tier, int skip = 0,
    TileTot* __restrict__ tile_tot = nullptr, uint32_t tile_slots = 1)
{
    // skip > 0: the complex already holds the arrangement of the first `skip` active functions
    // tile_tot (nullable): per-tile output totals of the streaming filter; `a` is then a tile-local slot
    const uint4 tv = __ldg(&tets[act_tet[a]]);
    if (!skip) cx.init();
    int seen = 0;
    for (int w = 0; w < W; ++w) {
        uint32_t mm = act_mask[(size_t)w * cap + a];
        while (mm) {
            int f = w * 32 + __ffs(mm) - 1;
            mm &= mm - 1;
            if (seen++ < skip) continue;
            double pv[4];
            pv[0] = __ldg(&vals[(size_t)f * V + tv.x]);
            pv[1] = __ldg(&vals[(size_t)f * V + tv.y]);
            pv[2] = __ldg(&vals[(size_t)f * V + tv.z]);
            pv[3] = __ldg(&vals[(size_t)f * V + tv.w]);
            cx.insert(pv);
        }
    }
    IsoScan<Caps> iso;
    if (!cx.err) {
        iso.run(cx);
        if (iso.nvi > 255 || iso.nfi > 255 || iso.nfv > 65535 || cx.nf > 65535) cx.err = 1;
    }
    if (cx.n_exact) atomicAdd(&gc->n_exact, cx.n_exact);
    if (cx.err == 1 && !last_tier) return false;
    if (cx.err) {
        if (atomicCAS(&gc->err, 0, cx.err == 1 ? RIN_ERR_CAPACITY : RIN_ERR_ARRANGEMENT) == 0)
            gc->err_tet = act_tet[a];
        rec_ref[a] = REF_GENERAL; // offset 0: the arena starts with an empty record
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
    if (tile_tot) {
        TileTot* tt = tile_tot + a / tile_slots;
        if (iso.nvi) atomicAdd(&tt->cand, (unsigned)iso.nvi);
        if (iso.nfi) atomicAdd(&tt->face, (unsigned)iso.nfi);
        if (iso.nfv) atomicAdd(&tt->fv, (unsigned)iso.nfv);
        if (iso.nbnd) atomicAdd(&gc->n_bnd_faces, (unsigned)iso.nbnd);
    }
    return true;
}

// Warp-cooperative version of general_ia_one (small tier): all 32 lanes call it with the same arguments.
template <class Caps, int W>
__device__ bool general_ia_one_warp(IAComplex<Caps>& cx, IAWarpScratch<Caps>& sc, uint32_t a,
    const uint4* __restrict__ tets, const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask,
    uint32_t cap, const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc, bool last_tier, int skip, int lane,
    TileTot* __restrict__ tile_tot = nullptr, uint32_t tile_slots = 1)
{
    const uint4 tv = __ldg(&tets[act_tet[a]]);
    if (!skip) {
        if (lane == 0) cx.init();
        __syncwarp();
    }
    int seen = 0;
    for (int w = 0; w < W; ++w) {
        uint32_t mm = act_mask[(size_t)w * cap + a];
        while (mm) {
            const int f = w * 32 + __ffs(mm) - 1;
            mm &= mm - 1;
            if (seen++ < skip) continue;
            double pv[4];
            pv[0] = __ldg(&vals[(size_t)f * V + tv.x]);
            pv[1] = __ldg(&vals[(size_t)f * V + tv.y]);
            pv[2] = __ldg(&vals[(size_t)f * V + tv.z]);
            pv[3] = __ldg(&vals[(size_t)f * V + tv.w]);
            warp_insert(cx, sc, pv, lane);
        }
    }
    __syncwarp();
    int err = cx.err;
    WarpIso<Caps> iso;
    if (!err) {
        iso.run(cx, lane);
        if (iso.nvi > 255 || iso.nfi > 255 || iso.nfv > 65535 || cx.nf > 65535) err = 1;
    }
    if (lane == 0 && cx.n_exact) atomicAdd(&gc->n_exact, cx.n_exact);
    if (err == 1 && !last_tier) return false;
    if (err) {
        if (lane == 0) {
            if (atomicCAS(&gc->err, 0, err == 1 ? RIN_ERR_CAPACITY : RIN_ERR_ARRANGEMENT) == 0)
                gc->err_tet = act_tet[a];
            rec_ref[a] = REF_GENERAL; // offset 0: the arena starts with an empty record
        }
        return true;
    }
    const uint32_t szal = iso.size_bytes();
    uint32_t off = 0;
    if (lane == 0) off = atomicAdd(&gc->arena_top, szal);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + szal > arena_cap) {
        if (lane == 0) {
            gc->arena_overflow = 1;
            rec_ref[a] = REF_GENERAL;
        }
        return true;
    }
    iso.write(cx, reinterpret_cast<uint32_t*>(arena + off), lane);
    if (lane == 0) {
        rec_ref[a] = REF_GENERAL | (off >> 2);
        if (tile_tot) {
            TileTot* tt = tile_tot + a / tile_slots;
            if (iso.nvi) atomicAdd(&tt->cand, (unsigned)iso.nvi);
            if (iso.nfi) atomicAdd(&tt->face, (unsigned)iso.nfi);
            if (iso.nfv) atomicAdd(&tt->fv, (unsigned)iso.nfv);
            if (iso.nbnd) atomicAdd(&gc->n_bnd_faces, (unsigned)iso.nbnd);
        }
    }
    return true;
}

// Small tier: one tet per warp, complex in shared memory, warp-cooperative insertion
// (ia_complex_warp.cuh): the stage is latency-bound (a few hundred general tets), so the serial
// dependency chain per tet is what sets its duration.
constexpr int GEN_SMALL_WARPS = 4;
constexpr int GEN_THREADS = 64;
// shared memory of one warp of the small tier: the complex and the re-packing scratch
struct alignas(16) SmallSlot
{
    IAComplex<IACapsSmall> cx;
    IAWarpScratch<IACapsSmall> sc;
};
// cx2 / lut2cx (nullable): complete 2-plane complexes per key.  A tet with >= 3 functions whose first
// two functions have a tabulated key starts from that complex (copied by the whole warp) and only
// inserts the remaining planes: every branch of the insertion depends on the vertex signs alone,
// which the key determines, so the state equals what two insertions would have produced.
template <int W>
__global__ void __launch_bounds__(GEN_SMALL_WARPS * 32) general_ia_small_kernel(
    const uint4* __restrict__ tets, const uint32_t* __restrict__ act_tet,
    const uint32_t* __restrict__ act_mask, uint32_t cap, const uint32_t* __restrict__ small_list,
    uint32_t* __restrict__ ovf_list, const IAComplex<IACapsSmall>* __restrict__ cx2,
    const uint16_t* __restrict__ lut2cx, const double* __restrict__ vals, uint32_t V,
    uint8_t* __restrict__ arena, uint32_t arena_cap, uint32_t* __restrict__ rec_ref,
    GeneralCounters* __restrict__ gc, TileTot* __restrict__ tile_tot = nullptr, uint32_t tile_slots = 1,
    uint32_t list_cap = 0xffffffffu)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    SmallSlot* s_slot = reinterpret_cast<SmallSlot*>(s_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = min(gc->n_small, list_cap); // a list that outgrew its capacity is cut (the pass is repeated)
    IAComplex<IACapsSmall>& cx = s_slot[warp].cx;
    IAWarpScratch<IACapsSmall>& sc = s_slot[warp].sc;
    for (uint32_t g = blockIdx.x * GEN_SMALL_WARPS + warp; g < n; g += gridDim.x * GEN_SMALL_WARPS) {
        const uint32_t a = small_list[g];
        int skip = 0;
        if (cx2) {
            // lane 0 derives the key of the first two active functions
            int entry = -1;
            double p0[4], p1[4];
            if (lane == 0) {
                uint32_t m[W];
                int kk = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    m[w] = act_mask[(size_t)w * cap + a];
                    kk += __popc(m[w]);
                }
                if (kk >= 3) {
                    const uint4 tv = __ldg(&tets[act_tet[a]]);
                    const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
                    const int f0 = nth_set_bit(m, W, 0), f1 = nth_set_bit(m, W, 1);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        p0[c] = __ldg(&vals[(size_t)f0 * V + vv[c]]);
                        p1[c] = __ldg(&vals[(size_t)f1 * V + vv[c]]);
                    }
                    unsigned ex = 0;
                    const int key = ia2_key(p0, p1, &ex);
                    if (ex) atomicAdd(&gc->n_exact, ex);
                    if (key >= 0 && lut2cx[key] != LUT_MISS) entry = lut2cx[key];
                }
            }
            entry = __shfl_sync(0xffffffffu, entry, 0);
            if (entry >= 0) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(cx2 + entry);
                uint32_t* dst = reinterpret_cast<uint32_t*>(&cx);
                for (int i = lane; i < (int)(sizeof(IAComplex<IACapsSmall>) / 4); i += 32) dst[i] = src[i];
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        cx.plv[0][c] = p0[c];
                        cx.plv[1][c] = p1[c];
                    }
                    cx.n_exact = 0;
                    cx.err = 0;
                }
                skip = 2;
            }
        }
        __syncwarp();
        if (!general_ia_one_warp<IACapsSmall, W>(cx, sc, a, tets, act_tet, act_mask, cap, vals, V, arena,
                arena_cap, rec_ref, gc, false, skip, lane, tile_tot, tile_slots)) {
            if (lane == 0) ovf_list[atomicAdd(&gc->n_ovf, 1u)] = a;
        }
        __syncwarp();
    }
}

// dumps the complete 2-plane complexes of the chosen witnesses (table generation)
__global__ void __launch_bounds__(GEN_THREADS) dump_ia2_kernel(const uint32_t* __restrict__ witness, uint32_t n,
    const double* __restrict__ vals, uint32_t V, IAComplex<IACapsSmall>* __restrict__ out, int* __restrict__ err)
{
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        const uint32_t w = witness[g]; // witness tet w owns vertices 4w..4w+3
        IAComplex<IACapsSmall> cx;
        cx.init();
        for (int f = 0; f < 2; ++f) {
            double pv[4];
            for (int c = 0; c < 4; ++c) pv[c] = vals[(size_t)f * V + 4 * w + c];
            cx.insert(pv);
        }
        if (cx.err) *err = cx.err;
        cx.n_exact = 0;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&cx);
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + g);
        for (int i = 0; i < (int)(sizeof(IAComplex<IACapsSmall>) / 4); ++i) dst[i] = src[i];
    }
}

// one tet through the serial big tier (complex in the calling thread's local memory)
template <int W>
__device__ __noinline__ void general_ia_big_one(uint32_t a, const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc, TileTot* __restrict__ tile_tot,
    uint32_t tile_slots)
{
    IAComplex<IACaps> cx;
    general_ia_one<IACaps, W>(cx, a, tets, act_tet, act_mask, cap, vals, V, arena, arena_cap, rec_ref, gc, true, 0,
        tile_tot, tile_slots);
}

// ---------------------------------------------------------------------------------------------
// Exclusive scan of the per-tile totals by ONE block (any multiple of 32 threads): every thread owns a run
// of consecutive tiles.  Grand totals -> *totals; capacities are checked here so that the kernels behind
// never write out of bounds: on overflow they do nothing and the host repeats the pass with larger buffers.
// ---------------------------------------------------------------------------------------------
struct PassTotals
{
    unsigned n_active, n_funcs, n_cand, n_faces, n_fv;
};
enum : unsigned {
    OVF_LIST = 1u, // a general work list outgrew its capacity
    OVF_ACT = 2u,  // active tets
    OVF_CAND = 4u, // vertex candidates
    OVF_FACE = 8u, // faces
    OVF_FV = 16u,  // face-vertex entries
    OVF_TABLE = 32u, // vertex hash table full
};
struct TileScanArgs
{
    const TileTot* tot;
    uint32_t n_tiles;
    TileTot* off; // n_tiles + 1 entries
    PassTotals* totals;
    uint32_t act_cap, cand_cap, face_cap, fv_cap;
    unsigned* overflow;
};

__device__ void tile_scan_block(const TileScanArgs& S)
{
    // warp w owns a contiguous run of tiles and reads it 32 tiles at a time (coalesced)
    __shared__ unsigned s_w[32][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t chunk = ((S.n_tiles + nwarps - 1) / nwarps + 31u) & ~31u;
    const uint32_t b = min(S.n_tiles, warp * chunk), e = min(S.n_tiles, b + chunk);
    // pass 1: the warp's totals
    unsigned c[5] = {0, 0, 0, 0, 0};
    for (uint32_t t0 = b + lane; t0 < e; t0 += 32 * 4) {
        uint4 a[4];
        unsigned f[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t t = t0 + 32 * q;
            const bool in = t < e;
            a[q] = in ? __ldcg(reinterpret_cast<const uint4*>(S.tot + t)) : make_uint4(0, 0, 0, 0);
            f[q] = in ? __ldcg(&S.tot[t].fv) : 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            c[0] += a[q].x;
            c[1] += a[q].y;
            c[2] += a[q].z;
            c[3] += a[q].w;
            c[4] += f[q];
        }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        for (int o = 16; o; o >>= 1) c[q] += __shfl_xor_sync(0xffffffffu, c[q], o);
        if (lane == 0) s_w[warp][q] = c[q];
    }
    __syncthreads();
    unsigned run[5] = {0, 0, 0, 0, 0}, total[5] = {0, 0, 0, 0, 0};
    for (int w = 0; w < nwarps; ++w)
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const unsigned x = s_w[w][q];
            if (w < warp) run[q] += x;
            total[q] += x;
        }
    if (threadIdx.x == 0) {
        S.totals->n_active = total[0];
        S.totals->n_funcs = total[1];
        S.totals->n_cand = total[2];
        S.totals->n_faces = total[3];
        S.totals->n_fv = total[4];
        unsigned o = 0;
        if (total[0] > S.act_cap) o |= OVF_ACT;
        if (total[2] > S.cand_cap) o |= OVF_CAND;
        if (total[3] > S.face_cap) o |= OVF_FACE;
        if (total[4] > S.fv_cap) o |= OVF_FV;
        if (o) atomicOr(S.overflow, o);
        uint4* o4 = reinterpret_cast<uint4*>(S.off + S.n_tiles);
        o4[0] = make_uint4(total[0], total[1], total[2], total[3]);
        o4[1] = make_uint4(total[4], 0, 0, 0);
    }
    // pass 2: exclusive prefixes, 32 tiles per round (the next round's loads are issued before this round's scan)
    uint4 a = make_uint4(0, 0, 0, 0);
    unsigned f = 0;
    if (b + lane < e) {
        a = __ldcg(reinterpret_cast<const uint4*>(S.tot + b + lane));
        f = __ldcg(&S.tot[b + lane].fv);
    }
    for (uint32_t t0 = b; t0 < e; t0 += 32) {
        const uint32_t t = t0 + lane, tn = t + 32;
        uint4 an = make_uint4(0, 0, 0, 0);
        unsigned fn = 0;
        if (tn < e) {
            an = __ldcg(reinterpret_cast<const uint4*>(S.tot + tn));
            fn = __ldcg(&S.tot[tn].fv);
        }
        unsigned v[5] = {a.x, a.y, a.z, a.w, f}, x[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            x[q] = v[q];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, x[q], o);
                if (lane >= o) x[q] += y;
            }
        }
        if (t < e) {
            uint4* o4 = reinterpret_cast<uint4*>(S.off + t);
            o4[0] = make_uint4(run[0] + x[0] - v[0], run[1] + x[1] - v[1], run[2] + x[2] - v[2], run[3] + x[3] - v[3]);
            o4[1] = make_uint4(run[4] + x[4] - v[4], 0, 0, 0);
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) run[q] += __shfl_sync(0xffffffffu, x[q], 31);
        a = an;
        f = fn;
    }
}

__global__ void __launch_bounds__(1024) scan_tiles5_kernel(const TileScanArgs S)
{
    tile_scan_block(S);
}

// Mid tier: the big list (more functions than the small tier takes) and the small tier's overflows, one tet
// per warp, complex in shared memory, warp-cooperative insertion.  Capacity overflow -> ovf2 list.
constexpr int GEN_MID_WARPS = 4;
struct alignas(16) MidSlot
{
    IAComplex<IACapsMid> cx;
    IAWarpScratch<IACapsMid> sc;
};
template <int W>
__global__ void __launch_bounds__(GEN_MID_WARPS * 32) general_ia_mid_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ ovf_list, uint32_t* __restrict__ ovf2_list,
    const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc, TileTot* __restrict__ tile_tot,
    uint32_t tile_slots, uint32_t list_cap, const TileScanArgs S)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    MidSlot* s_slot = reinterpret_cast<MidSlot*>(s_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nb = min(gc->n_big, list_cap), n = nb + min(gc->n_ovf, list_cap);
    bool worked = false;
    for (uint32_t g = blockIdx.x * GEN_MID_WARPS + warp; g < n; g += gridDim.x * GEN_MID_WARPS) {
        const uint32_t a = g < nb ? big_list[g] : ovf_list[g - nb];
        worked = true;
        __syncwarp();
        if (!general_ia_one_warp<IACapsMid, W>(s_slot[warp].cx, s_slot[warp].sc, a, tets, act_tet, act_mask, cap,
                vals, V, arena, arena_cap, rec_ref, gc, false, 0, lane, tile_tot, tile_slots)) {
            if (ovf2_list) {
                if (lane == 0) ovf2_list[atomicAdd(&gc->n_ovf2, 1u)] = a;
            } else if (lane == 0) {
                // what this tier cannot hold (rare): lane 0 runs the serial big tier in local memory
                general_ia_big_one<W>(a, tets, act_tet, act_mask, cap, vals, V, arena, arena_cap, rec_ref, gc,
                    tile_tot, tile_slots);
            }
        }
        __syncwarp();
    }
    // The last block to finish scans the tile totals: all general records exist by then.  Only blocks that
    // produced records need the fence before they sign off.
    if (S.tot) {
        __shared__ bool s_scan;
        if (worked) __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_scan = atomicAdd(&gc->done_blocks, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_scan) {
            __threadfence();
            tile_scan_block(S);
        }
    }
}

// Big tier as a kernel of its own (table generation): complexes in per-thread local memory.
template <int W>
__global__ void __launch_bounds__(GEN_THREADS) general_ia_big_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    const uint32_t* __restrict__ list, const unsigned* __restrict__ n_list,
    const double* __restrict__ vals, uint32_t V, uint8_t* __restrict__ arena, uint32_t arena_cap,
    uint32_t* __restrict__ rec_ref, GeneralCounters* __restrict__ gc)
{
    const uint32_t n = *n_list;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
        general_ia_big_one<W>(list[g], tets, act_tet, act_mask, cap, vals, V, arena, arena_cap, rec_ref, gc, nullptr, 1);
}

// ---------------------------------------------------------------------------------------------
// K5a: per-active-tet output counts + exclusive scan (4 lanes: iso verts, iso faces, face-vertex
// entries, active functions).  offs[a] = (vert offset, face offset, face-vertex offset, CRS offset)
// ---------------------------------------------------------------------------------------------
struct ScanTotals
{
    unsigned tile_counter;
    unsigned n_cand, n_faces, n_fv, n_funcs;
};

__device__ __forceinline__ const uint8_t* record_ptr(uint32_t ref, const uint8_t* lut_blob, const uint8_t* arena)
{
    return (ref & REF_GENERAL) ? arena + (size_t)(ref & ~REF_FLAGS) * 4 : lut_blob + (size_t)ref * 4;
}

constexpr int CS_ITEMS = 4;
constexpr int CS_TILE = 256 * CS_ITEMS;

template <int W>
__global__ void __launch_bounds__(256) count_scan_kernel(const uint32_t* __restrict__ rec_ref,
    const uint32_t* __restrict__ act_mask, uint32_t cap, uint32_t n_active,
    const uint8_t* __restrict__ lut_blob, const uint8_t* __restrict__ arena, uint4* __restrict__ offs,
    volatile unsigned long long* __restrict__ statusA, volatile unsigned long long* __restrict__ statusB,
    ScanTotals* __restrict__ tot, const unsigned* __restrict__ n_dev = nullptr)
{
    __shared__ unsigned s_tile;
    __shared__ uint4 s_warp[8];
    __shared__ uint4 s_base;
    if (n_dev) n_active = min(n_active, *n_dev); // launched for the capacity: the count is still on the device
    if (threadIdx.x == 0) s_tile = atomicAdd(&tot->tile_counter, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if ((size_t)tile * CS_TILE >= n_active) return; // beyond the count: no later tile looks back at this one
    const uint32_t a0 = tile * CS_TILE + threadIdx.x * CS_ITEMS; // CS_ITEMS consecutive active tets per thread
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 ci[CS_ITEMS];
    uint4 c = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < CS_ITEMS; ++j) {
        const uint32_t a = a0 + j;
        ci[j] = make_uint4(0, 0, 0, 0);
        if (a < n_active) {
            const uint8_t* r = record_ptr(rec_ref[a], lut_blob, arena);
            const uint32_t h = *reinterpret_cast<const uint32_t*>(r);
            ci[j].x = h & 255;
            ci[j].y = (h >> 8) & 255;
            ci[j].z = h >> 16;
            for (int w = 0; w < W; ++w) ci[j].w += __popc(act_mask[(size_t)w * cap + a]);
        }
        c.x += ci[j].x;
        c.y += ci[j].y;
        c.z += ci[j].z;
        c.w += ci[j].w;
    }
    uint4 x = c;
    for (int o = 1; o < 32; o <<= 1) {
        uint4 y;
        y.x = __shfl_up_sync(0xffffffffu, x.x, o);
        y.y = __shfl_up_sync(0xffffffffu, x.y, o);
        y.z = __shfl_up_sync(0xffffffffu, x.z, o);
        y.w = __shfl_up_sync(0xffffffffu, x.w, o);
        if (lane >= o) {
            x.x += y.x;
            x.y += y.y;
            x.z += y.z;
            x.w += y.w;
        }
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        // exclusive prefix of the 8 warp totals (lanes 0..7), block total in lane 7
        uint4 t = (lane < 8) ? s_warp[lane] : make_uint4(0, 0, 0, 0), x8 = t;
        for (int o = 1; o < 8; o <<= 1) {
            uint4 y;
            y.x = __shfl_up_sync(0xffffffffu, x8.x, o);
            y.y = __shfl_up_sync(0xffffffffu, x8.y, o);
            y.z = __shfl_up_sync(0xffffffffu, x8.z, o);
            y.w = __shfl_up_sync(0xffffffffu, x8.w, o);
            if (lane >= o) {
                x8.x += y.x;
                x8.y += y.y;
                x8.z += y.z;
                x8.w += y.w;
            }
        }
        if (lane < 8) s_warp[lane] = make_uint4(x8.x - t.x, x8.y - t.y, x8.z - t.z, x8.w - t.w);
        uint4 run;
        run.x = __shfl_sync(0xffffffffu, x8.x, 7);
        run.y = __shfl_sync(0xffffffffu, x8.y, 7);
        run.z = __shfl_sync(0xffffffffu, x8.z, 7);
        run.w = __shfl_sync(0xffffffffu, x8.w, 7);
        uint32_t e0, e1, e2, e3;
        tile_lookback_warp(statusA, (int)tile, run.x, run.y, e0, e1);
        tile_lookback_warp(statusB, (int)tile, run.z, run.w, e2, e3);
        if (lane == 0) {
            s_base = make_uint4(e0, e1, e2, e3);
            if (tile == (n_active + CS_TILE - 1) / CS_TILE - 1) {
                tot->n_cand = e0 + run.x;
                tot->n_faces = e1 + run.y;
                tot->n_fv = e2 + run.z;
                tot->n_funcs = e3 + run.w;
            }
        }
    }
    __syncthreads();
    {
        const uint4 b = s_base, wv = s_warp[warp];
        uint4 run = make_uint4(b.x + wv.x + x.x - c.x, b.y + wv.y + x.y - c.y, b.z + wv.z + x.z - c.z,
            b.w + wv.w + x.w - c.w);
#pragma unroll
        for (int j = 0; j < CS_ITEMS; ++j) {
            if (a0 + j < n_active) offs[a0 + j] = run;
            run.x += ci[j].x;
            run.y += ci[j].y;
            run.z += ci[j].z;
            run.w += ci[j].w;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5b: emit vertex candidates (canonical keys) and face records.
// Replaces the body of extract_iso_mesh (src/extract_mesh.cpp:93-261): classification of an iso
// vertex by the number of simplex-boundary planes among its three planes and the keys
//   on tet vertex (v)                :197-225        on tet edge (vmin, vmax, fn)       :113-151
//   on tet face (v0<=v1<=v2, fa, fb) :153-183        interior: never shared             :185-195
// cand_key  = (v0, v1, v2, fa | fb<<16) with unused slots 0xffffffff / 0xffff
// cand_pay  = (tet, local | simplex_size<<8 | dedup<<16, f0 | f1<<16, f2)
// face_hdr  = (tet, local face | n<<16 | boundary<<24, function id, face-vertex offset)
// ---------------------------------------------------------------------------------------------
template <int W, bool SMEM_BLOB>
__global__ void __launch_bounds__(256) emit_ia_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, const uint32_t* __restrict__ act_mask, uint32_t cap,
    uint32_t n_active, const uint32_t* __restrict__ rec_ref, const uint4* __restrict__ offs,
    const uint8_t* __restrict__ lut_blob, uint32_t lut_bytes, const uint8_t* __restrict__ arena,
    uint4* __restrict__ cand_key, uint4* __restrict__ cand_pay, uint4* __restrict__ face_hdr,
    uint32_t* __restrict__ fv_ref, unsigned* __restrict__ n_bndry_faces)
{
    extern __shared__ __align__(16) uint8_t s_blob[];
    if (SMEM_BLOB) {
        for (uint32_t i = threadIdx.x; i < lut_bytes / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(s_blob)[i] = reinterpret_cast<const uint32_t*>(lut_blob)[i];
        __syncthreads();
    }
    const uint32_t* table = SMEM_BLOB ? reinterpret_cast<const uint32_t*>(s_blob)
                                      : reinterpret_cast<const uint32_t*>(lut_blob);
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const uint32_t ref = rec_ref[a];
        const uint32_t* r = (ref & REF_GENERAL)
                                ? reinterpret_cast<const uint32_t*>(arena) + (size_t)(ref & ~REF_GENERAL)
                                : table + ref;
        const uint32_t hdr = r[0];
        const int nv = hdr & 255, nf = (hdr >> 8) & 255;
        if (nv == 0 && nf == 0) continue;
        const uint32_t t = act_tet[a];
        const uint4 tv4 = __ldg(&tets[t]);
        const uint32_t tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
        uint32_t m[W];
#pragma unroll
        for (int w = 0; w < W; ++w) m[w] = act_mask[(size_t)w * cap + a];
        // the first four active functions in registers (tables never need more)
        uint32_t fl[4] = {0xffffu, 0xffffu, 0xffffu, 0xffffu};
        {
            int q = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t mm = m[w];
                while (mm && q < 4) {
                    fl[q++] = w * 32 + __ffs(mm) - 1;
                    mm &= mm - 1;
                }
            }
        }
        auto func_of = [&](int pl) -> uint32_t { // plane id (>= 4) -> global function id
            const int j = pl - 4;
            return j < 4 ? fl[j] : (uint32_t)nth_set_bit(m, W, j);
        };
        const uint4 o = offs[a];
        const uint32_t* p = r + 1;
        for (int i = 0; i < nv; ++i) {
            const uint32_t e = p[i];
            const int local = e & 255;
            const int p0 = (e >> 8) & 255, p1 = (e >> 16) & 255, p2 = e >> 24;
            // planes ascend: boundary planes (< 4) come first
            const int nb = (p0 < 4) + (p1 < 4) + (p2 < 4);
            uint4 key, pay;
            pay.x = t;
            uint32_t f0 = 0xffffu, f1 = 0xffffu, f2 = 0xffffu;
            if (nb == 0) {
                f0 = func_of(p0);
                f1 = func_of(p1);
                f2 = func_of(p2);
                key = make_uint4(tv[0], tv[1], tv[2], tv[3]); // unused: interior vertices are unique
                pay.y = (uint32_t)local | (4u << 8);
            } else {
                uint32_t c0, c1 = NONE32, c2 = NONE32;
                int ncn;
                if (nb == 3) { // on the tet corner not listed
                    c0 = tv[6 - p0 - p1 - p2];
                    ncn = 1;
                } else if (nb == 2) { // on the tet edge between the two corners not listed
                    f0 = func_of(p2);
                    const unsigned rest = 0xfu & ~((1u << p0) | (1u << p1));
                    const int a0 = __ffs(rest) - 1, a1 = 31 - __clz(rest);
                    c0 = min(tv[a0], tv[a1]);
                    c1 = max(tv[a0], tv[a1]);
                    ncn = 2;
                } else { // on the tet face opposite corner p0
                    f0 = func_of(p1);
                    f1 = func_of(p2);
                    const uint32_t x = tv[p0 == 0 ? 1 : 0], y = tv[p0 <= 1 ? 2 : 1], z = tv[p0 == 3 ? 2 : 3];
                    const uint32_t lo = min(x, min(y, z)), hi = max(x, max(y, z));
                    c0 = lo;
                    c1 = x ^ y ^ z ^ lo ^ hi; // the middle one
                    c2 = hi;
                    ncn = 3;
                }
                key = make_uint4(c0, c1, c2, f0 | (f1 << 16)); // function ids in plane order (:138,:167-168)
                pay.y = (uint32_t)local | ((uint32_t)ncn << 8) | (1u << 16);
            }
            pay.z = f0 | (f1 << 16);
            pay.w = f2;
            cand_key[o.x + i] = key;
            cand_pay[o.x + i] = pay;
        }
        p += nv;
        uint32_t fvo = o.z;
        unsigned nbf = 0;
        for (int j = 0; j < nf; ++j) {
            const uint32_t e = *p++;
            const uint32_t local = e & 0xffffu;
            const int sp = (e >> 16) & 255, n = (e >> 24) & 127, bnd = e >> 31;
            // func_index.first = func_in_tet[supporting_plane - 4 + start] (:249,:258); for a face
            // coplanar with a tet face sp < 4 and the reference's index wraps to an earlier CRS entry
            uint32_t f;
            if (sp > 3)
                f = func_of(sp);
            else {
                // QUIRK kept from the reference: CRS entry (start + sp - 4) belongs to an earlier tet
                f = NONE32;
                int back = 4 - sp; // how many CRS entries before this tet's start
                for (uint32_t ap = a; ap > 0 && back > 0;) {
                    --ap;
                    uint32_t mp[W];
                    int kp = 0;
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        mp[w] = act_mask[(size_t)w * cap + ap];
                        kp += __popc(mp[w]);
                    }
                    if (back <= kp) {
                        f = (uint32_t)nth_set_bit(mp, W, kp - back);
                        back = 0;
                    } else
                        back -= kp;
                }
            }
            face_hdr[o.y + j] = make_uint4(t, local | ((uint32_t)n << 16) | ((uint32_t)bnd << 24), f, fvo);
            for (int k0 = 0; k0 < n; k0 += 4) {
                const uint32_t x = *p++;
                for (int k = k0; k < n && k < k0 + 4; ++k) fv_ref[fvo + k] = o.x + ((x >> (8 * (k - k0))) & 255);
            }
            fvo += n;
            nbf += bnd;
        }
        if (nbf) atomicAdd(n_bndry_faces, nbf);
    }
}

// ---------------------------------------------------------------------------------------------
// K6: deduplication with first-occurrence numbering.
// The reference's try_emplace(key, next id) (src/extract_mesh.cpp:139,169,214,244) gives a shared
// vertex the id of its FIRST candidate in (tet, local index) order.  Here: every candidate
// atomicMin's its index into an open-addressing table slot owned by its key -> the slot ends up
// holding the first candidate (order independent, hence deterministic); representatives are
// flagged and ranked by an exclusive scan in candidate order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hash_insert_kernel(const uint4* __restrict__ keys,
    const uint4* __restrict__ pay, uint32_t n, uint32_t* __restrict__ table, uint32_t mask,
    uint32_t* __restrict__ slot_of)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        if (!((pay[c].y >> 16) & 1)) {
            slot_of[c] = NONE32;
            continue;
        }
        const uint4 k = keys[c];
        uint32_t h = hash4(k) & mask;
        for (;;) {
            uint32_t cur = table[h];
            if (cur == NONE32) {
                cur = atomicCAS(&table[h], NONE32, c);
                if (cur == NONE32) break;
            }
            if (key_eq(keys[cur], k)) {
                atomicMin(&table[h], c);
                break;
            }
            h = (h + 1) & mask;
        }
        slot_of[c] = h;
    }
}

// rep[c] = first candidate with the same key; single-lane look-back scan of the representative
// flags gives vid[c] (valid at representatives).
constexpr uint32_t CAND_DEDUP = 1u << 16, CAND_DEAD = 1u << 17; // bits of cand_pay.y
constexpr uint32_t FACE_BND = 1u << 24, FACE_INACTIVE = 1u << 25; // bits of face_hdr.y

__global__ void __launch_bounds__(256) rank_reps_kernel(const uint32_t* __restrict__ table,
    const uint32_t* __restrict__ slot_of, const uint4* __restrict__ pay /* nullable: marks dead candidates */,
    uint32_t n, uint32_t* __restrict__ rep, uint32_t* __restrict__ vid,
    volatile unsigned long long* __restrict__ status, unsigned* __restrict__ tile_counter,
    unsigned* __restrict__ n_unique)
{
    __shared__ unsigned s_tile, s_base;
    __shared__ unsigned s_warp[8];
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ITEMS = 4;
    const uint32_t base = tile * 256 * ITEMS + threadIdx.x * ITEMS;
    uint32_t r[ITEMS];
    unsigned cnt = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t c = base + j;
        r[j] = NONE32;
        if (c < n) {
            const uint32_t s = slot_of[c];
            r[j] = (s == NONE32) ? c : table[s];
            if (pay && (pay[c].y & CAND_DEAD)) r[j] = NONE32; // reserved slot that was never activated
            cnt += (r[j] == c);
        }
    }
    unsigned x = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned t = (lane < 8) ? s_warp[lane] : 0, x8 = t;
        for (int o = 1; o < 8; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x8, o);
            if (lane >= o) x8 += y;
        }
        if (lane < 8) s_warp[lane] = x8 - t;
        const unsigned run = __shfl_sync(0xffffffffu, x8, 7);
        uint32_t e0, e1;
        tile_lookback_warp(status, (int)tile, run, 0, e0, e1);
        if (lane == 0) {
            s_base = e0;
            if (tile == (n + 256 * ITEMS - 1) / (256 * ITEMS) - 1) *n_unique = e0 + run;
        }
    }
    __syncthreads();
    unsigned id = s_base + s_warp[warp] + x - cnt;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t c = base + j;
        if (c < n) {
            rep[c] = r[j];
            vid[c] = (r[j] == c) ? id++ : NONE32;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K7: write the unique vertices (IsoVert fields) and their coordinates.
// compute_iso_vert_xyz (src/extract_mesh.cpp:1446-1538) with compute_barycentric_coords
// (src/extract_mesh.h:111-159): same operation order, no FMA contraction (-fmad=false).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) write_verts_ia_kernel(const uint4* __restrict__ cand_key,
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
        uint32_t fi[3] = {pay.z & 0xffffu, pay.z >> 16, pay.w & 0xffffu};
        for (int q = 0; q < 3; ++q)
            if (fi[q] == 0xffffu) fi[q] = NONE32;
        v_tet[id] = pay.x;
        v_local[id] = (uint8_t)(pay.y & 255);
        v_size[id] = (uint8_t)size;
        v_simplex[id] = make_uint4(sv[0], sv[1], sv[2], sv[3]);
        v_funcs[id] = make_uint4(fi[0], fi[1], fi[2], NONE32);
        double out[3];
#define PT(v, c) pts[3 * (size_t)(v) + (c)]
#define FV(v, f) vals[(size_t)(f) * V + (v)]
        if (size == 1) {
            for (int d = 0; d < 3; ++d) out[d] = PT(sv[0], d);
        } else if (size == 2) {
            const double f1 = FV(sv[0], fi[0]), f2 = FV(sv[1], fi[0]);
            const double b0 = f2 / (f2 - f1), b1 = 1 - b0;
            for (int d = 0; d < 3; ++d) out[d] = b0 * PT(sv[0], d) + b1 * PT(sv[1], d);
        } else if (size == 3) {
            double p1[3], p2[3];
            for (int k = 0; k < 3; ++k) {
                p1[k] = FV(sv[k], fi[0]);
                p2[k] = FV(sv[k], fi[1]);
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
                p1[k] = FV(sv[k], fi[0]);
                p2[k] = FV(sv[k], fi[1]);
                p3[k] = FV(sv[k], fi[2]);
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

// face vertex references -> final vertex ids
__global__ void __launch_bounds__(256) remap_face_verts_kernel(const uint32_t* __restrict__ fv_ref,
    uint32_t n, const uint32_t* __restrict__ rep, const uint32_t* __restrict__ vid, uint32_t* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t r = rep[fv_ref[i]];
        out[i] = (r == NONE32) ? NONE32 : vid[r];
    }
}

// generic (no iso-face on a tet boundary): one output face per face record
__global__ void __launch_bounds__(256) write_faces_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
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
        f_funcs[2 * (size_t)i] = h.z;
        f_funcs[2 * (size_t)i + 1] = NONE32;
    }
}

// ---------------------------------------------------------------------------------------------
// Cell-grouping maps (the second extract_iso_mesh overload, src/extract_mesh.cpp:268-566): for every
// active tet, the global id of every local vertex of its complex (global_vId_of_tet_vert: iso-vertex id,
// tet corners encoded -(grid vertex id)-1, :402,518) and the iso-face id of every local face
// (iso_fId_of_tet_face, None when the face is not on an iso-surface, :529-556).
// Local vertices 0..3 of a complex are the tet corners and never move; every other vertex was created
// by an input plane and is an iso-vertex, so the record lists it.  The number of local faces is the
// record's trailing word.  Pass 1 counts, a one-block scan makes the CRS offsets, pass 2 writes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tetmap_count_kernel(const uint32_t* __restrict__ rec_ref,
    uint32_t n_active, const uint8_t* __restrict__ lut_blob, const uint8_t* __restrict__ arena,
    uint2* __restrict__ cnt)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(record_ptr(rec_ref[a], lut_blob, arena));
        const int nv = r[0] & 255, nf = (r[0] >> 8) & 255;
        int top = 4;
        for (int i = 0; i < nv; ++i) top = max(top, (int)(r[1 + i] & 255) + 1);
        const uint32_t* p = r + 1 + nv;
        for (int q = 0; q < nf; ++q) p += rec_face_words((p[0] >> 24) & 127);
        cnt[a] = make_uint2((uint32_t)top, p[0]);
    }
}

// single-block exclusive scan of pairs; off has n + 1 entries
__global__ void __launch_bounds__(1024) scan_pairs_kernel(const uint2* __restrict__ cnt, uint32_t n,
    uint2* __restrict__ off)
{
    __shared__ uint32_t s_w[32][2];
    __shared__ uint32_t s_run[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) s_run[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint2 c = i < n ? cnt[i] : make_uint2(0, 0);
        uint32_t x[2] = {c.x, c.y};
        for (int o = 1; o < 32; o <<= 1)
            for (int q = 0; q < 2; ++q) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x[q], o);
                if (lane >= o) x[q] += y;
            }
        if (lane == 31)
            for (int q = 0; q < 2; ++q) s_w[warp][q] = x[q];
        __syncthreads();
        if (warp == 0)
            for (int q = 0; q < 2; ++q) {
                const uint32_t t = s_w[lane][q];
                uint32_t y = t;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t z = __shfl_up_sync(0xffffffffu, y, o);
                    if (lane >= o) y += z;
                }
                s_w[lane][q] = y - t;
            }
        __syncthreads();
        if (i < n) off[i] = make_uint2(s_run[0] + s_w[warp][0] + x[0] - c.x, s_run[1] + s_w[warp][1] + x[1] - c.y);
        __syncthreads();
        if (threadIdx.x == 1023)
            for (int q = 0; q < 2; ++q) s_run[q] += s_w[31][q] + x[q];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = make_uint2(s_run[0], s_run[1]);
}

// frep / fpos: face slot -> representative slot -> final face id (degenerate runs only, else nullptr)
__global__ void __launch_bounds__(256) tetmap_write_kernel(const uint4* __restrict__ tets,
    const uint32_t* __restrict__ act_tet, uint32_t n_active, const uint32_t* __restrict__ rec_ref,
    const uint4* __restrict__ offs, const uint8_t* __restrict__ lut_blob, const uint8_t* __restrict__ arena,
    const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ table, const uint32_t* __restrict__ frep,
    const uint4* __restrict__ fpos, const uint2* __restrict__ off, long long* __restrict__ vmap,
    uint32_t* __restrict__ fmap)
{
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(record_ptr(rec_ref[a], lut_blob, arena));
        const int nv = r[0] & 255, nf = (r[0] >> 8) & 255;
        const uint4 tv4 = __ldg(&tets[act_tet[a]]);
        const uint32_t tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
        const uint4 o = offs[a];
        const uint2 b = off[a], e = off[a + 1];
        for (int j = 0; j < 4; ++j) vmap[b.x + j] = -(long long)tv[j] - 1;
        for (int i = 0; i < nv; ++i) {
            const int local = r[1 + i] & 255;
            if (local >= 4) vmap[b.x + local] = (long long)final_vid(o.x + i, slot_of, table);
        }
        for (uint32_t f = b.y; f < e.y; ++f) fmap[f] = NONE32;
        const uint32_t* p = r + 1 + nv;
        for (int q = 0; q < nf; ++q) {
            const uint32_t slot = o.y + q;
            fmap[b.y + (p[0] & 0xffffu)] = frep ? fpos[frep[slot]].x : slot;
            p += rec_face_words((p[0] >> 24) & 127);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Degenerate inputs only: iso-faces lying on a tet boundary are shared by two tets and are
// deduplicated by (smallest, second smallest, largest) vertex id; the first tet keeps the face and
// collects the (tet, local face) pairs of the others (src/extract_mesh.cpp:240-253,
// compute_iso_face_key src/extract_mesh.h:68-92).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bface_keys_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
    const uint32_t* __restrict__ fverts, uint4* __restrict__ fkeys)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 h = face_hdr[i];
        if (!((h.y >> 24) & 1)) continue;
        const int nv = (h.y >> 16) & 255;
        uint32_t mn = fverts[h.w], mx = mn;
        int mn_pos = 0;
        for (int k = 1; k < nv; ++k) {
            uint32_t v = fverts[h.w + k];
            if (v < mn) {
                mn = v;
                mn_pos = k;
            } else if (v > mx)
                mx = v;
        }
        uint32_t second = mx + 1;
        for (int k = 0; k < nv; ++k) {
            uint32_t v = fverts[h.w + k];
            if (k != mn_pos && v < second) second = v;
        }
        fkeys[i] = make_uint4(mn, second, mx, 0);
    }
}

__global__ void __launch_bounds__(256) bface_insert_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
    const uint4* __restrict__ fkeys, uint32_t* __restrict__ table, uint32_t mask, uint32_t* __restrict__ slot_of)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        if (!((face_hdr[c].y >> 24) & 1)) {
            slot_of[c] = NONE32;
            continue;
        }
        const uint4 k = fkeys[c];
        uint32_t h = hash4(k) & mask;
        for (;;) {
            uint32_t cur = table[h];
            if (cur == NONE32) {
                cur = atomicCAS(&table[h], NONE32, c);
                if (cur == NONE32) break;
            }
            if (key_eq(fkeys[cur], k)) {
                atomicMin(&table[h], c);
                break;
            }
            h = (h + 1) & mask;
        }
        slot_of[c] = h;
    }
}

// frep[i] = representative face; ndup[rep] counts the later duplicates
__global__ void __launch_bounds__(256) bface_reps_kernel(const uint32_t* __restrict__ table,
    const uint32_t* __restrict__ slot_of, uint32_t n, uint32_t* __restrict__ frep, uint32_t* __restrict__ ndup)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = slot_of[i];
        const uint32_t r = (s == NONE32) ? i : table[s];
        frep[i] = r;
        if (r != i) atomicAdd(&ndup[r], 1u);
    }
}

// single-block exclusive scan of (kept, kept ? nverts : 0, kept ? 1 + ndup : 0) -> pos[i]
__global__ void __launch_bounds__(1024) bface_scan_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
    const uint32_t* __restrict__ frep, const uint32_t* __restrict__ ndup, uint4* __restrict__ pos,
    uint32_t* __restrict__ totals)
{
    __shared__ uint32_t s_w[32][3];
    __shared__ uint32_t s_run[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3) s_run[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        uint32_t c[3] = {0, 0, 0};
        if (i < n && frep[i] == i) {
            c[0] = 1;
            c[1] = (face_hdr[i].y >> 16) & 255;
            c[2] = 1 + ndup[i];
        }
        uint32_t x[3] = {c[0], c[1], c[2]};
        for (int o = 1; o < 32; o <<= 1)
            for (int q = 0; q < 3; ++q) {
                uint32_t y = __shfl_up_sync(0xffffffffu, x[q], o);
                if (lane >= o) x[q] += y;
            }
        if (lane == 31)
            for (int q = 0; q < 3; ++q) s_w[warp][q] = x[q];
        __syncthreads();
        if (warp == 0) {
            for (int q = 0; q < 3; ++q) {
                uint32_t t = s_w[lane][q], y = t;
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t z = __shfl_up_sync(0xffffffffu, y, o);
                    if (lane >= o) y += z;
                }
                s_w[lane][q] = y - t;
            }
        }
        __syncthreads();
        if (i < n)
            pos[i] = make_uint4(s_run[0] + s_w[warp][0] + x[0] - c[0], s_run[1] + s_w[warp][1] + x[1] - c[1],
                s_run[2] + s_w[warp][2] + x[2] - c[2], 0);
        __syncthreads();
        if (threadIdx.x == 1023)
            for (int q = 0; q < 3; ++q) s_run[q] += s_w[31][q] + x[q];
        __syncthreads();
    }
    if (threadIdx.x < 3) totals[threadIdx.x] = s_run[threadIdx.x];
}

__global__ void __launch_bounds__(256) bface_write_kernel(const uint4* __restrict__ face_hdr, uint32_t n,
    const uint32_t* __restrict__ fverts_in, const uint32_t* __restrict__ frep, const uint4* __restrict__ pos,
    uint32_t* __restrict__ cursor, const uint32_t* __restrict__ totals, uint32_t* __restrict__ f_off,
    uint32_t* __restrict__ f_verts, uint32_t* __restrict__ f_toff, uint32_t* __restrict__ f_tets,
    uint32_t* __restrict__ f_funcs, int mi_labels)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        if (i == n) {
            f_off[totals[0]] = totals[1];
            f_toff[totals[0]] = totals[2];
            continue;
        }
        const uint4 h = face_hdr[i];
        const uint32_t r = frep[i];
        if (r == i) {
            const uint4 p = pos[i];
            const int nv = (h.y >> 16) & 255;
            f_off[p.x] = p.y;
            for (int k = 0; k < nv; ++k) f_verts[p.y + k] = fverts_in[h.w + k];
            f_toff[p.x] = p.z;
            f_tets[2 * (size_t)p.z] = h.x;
            f_tets[2 * (size_t)p.z + 1] = h.y & 0xffffu;
            if (mi_labels) {
                const uint32_t a = h.z & 0xffffu, b = h.z >> 16;
                f_funcs[2 * (size_t)p.x] = (a == 0xffffu) ? NONE32 : a;
                f_funcs[2 * (size_t)p.x + 1] = b;
            } else {
                f_funcs[2 * (size_t)p.x] = h.z;
                f_funcs[2 * (size_t)p.x + 1] = NONE32;
            }
        } else if (r != NONE32) {
            const uint4 p = pos[r];
            const uint32_t slot = p.z + 1 + atomicAdd(&cursor[r], 1u);
            f_tets[2 * (size_t)slot] = h.x;
            f_tets[2 * (size_t)slot + 1] = h.y & 0xffffu;
        }
    }
}

// duplicates were appended in arbitrary order: restore (tet, local face) order per face
__global__ void __launch_bounds__(256) bface_sort_pairs_kernel(uint32_t n_faces, const uint32_t* __restrict__ f_toff,
    uint32_t* __restrict__ f_tets)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_faces; j += gridDim.x * blockDim.x) {
        const uint32_t b = f_toff[j] + 1, e = f_toff[j + 1];
        for (uint32_t x = b + 1; x < e; ++x) {
            uint32_t t0 = f_tets[2 * (size_t)x], t1 = f_tets[2 * (size_t)x + 1];
            uint32_t y = x;
            while (y > b && (f_tets[2 * (size_t)(y - 1)] > t0 ||
                                (f_tets[2 * (size_t)(y - 1)] == t0 && f_tets[2 * (size_t)(y - 1) + 1] > t1))) {
                f_tets[2 * (size_t)y] = f_tets[2 * (size_t)(y - 1)];
                f_tets[2 * (size_t)y + 1] = f_tets[2 * (size_t)(y - 1) + 1];
                --y;
            }
            f_tets[2 * (size_t)y] = t0;
            f_tets[2 * (size_t)y + 1] = t1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Slab-boundary exchange (multi-GPU): vertices whose minimal simplex lies in the vertex range
// shared with another rank.
// ---------------------------------------------------------------------------------------------
// appends (key, id) of the unique vertices with simplex inside [lo, hi]; id = local id, or the own
// index when own_only (foreign vertices are skipped)
__global__ void __launch_bounds__(256) boundary_select_kernel(const uint4* __restrict__ v_key,
    const uint8_t* __restrict__ v_size, uint32_t n, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ own_idx,
    uint4* __restrict__ out_keys, uint32_t* __restrict__ out_ids, uint32_t cap, unsigned* __restrict__ n_out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int sz = v_size[i];
        if (sz >= 4) continue;
        const uint4 k = v_key[i];
        bool in = k.x >= lo && k.x <= hi;
        if (sz >= 2) in &= k.y >= lo && k.y <= hi;
        if (sz >= 3) in &= k.z >= lo && k.z <= hi;
        if (!in) continue;
        uint32_t id = i;
        if (own_idx) {
            id = own_idx[i];
            if (id == NONE32) continue;
        }
        const unsigned p = atomicAdd(n_out, 1u);
        if (p < cap) {
            out_keys[p] = k;
            out_ids[p] = id;
        }
    }
}

// open-addressing set of foreign keys: table[h] = index into fkeys
__global__ void __launch_bounds__(256) foreign_insert_kernel(const uint4* __restrict__ fkeys, uint32_t m,
    uint32_t* __restrict__ table, uint32_t mask)
{
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < m; c += gridDim.x * blockDim.x) {
        const uint4 k = fkeys[c];
        uint32_t h = hash4(k) & mask;
        for (;;) {
            uint32_t cur = atomicCAS(&table[h], NONE32, c);
            if (cur == NONE32 || key_eq(fkeys[cur], k)) break;
            h = (h + 1) & mask;
        }
    }
}

__device__ __forceinline__ uint32_t foreign_lookup(const uint4* fkeys, const uint32_t* table, uint32_t mask, uint4 k)
{
    uint32_t h = hash4(k) & mask;
    for (;;) {
        const uint32_t cur = table[h];
        if (cur == NONE32) return NONE32;
        if (key_eq(fkeys[cur], k)) return cur;
        h = (h + 1) & mask;
    }
}

// flag[i] = 1 when vertex i is owned by this rank (not found among the lower ranks' keys)
__global__ void __launch_bounds__(256) mark_foreign_kernel(const uint4* __restrict__ v_key,
    const uint8_t* __restrict__ v_size, uint32_t n, const uint4* __restrict__ fkeys,
    const uint32_t* __restrict__ table, uint32_t mask, uint32_t* __restrict__ own_flag, uint32_t v_lo)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t own = 1;
        if (v_size[i] < 4 && mask && foreign_lookup(fkeys, table, mask, v_key[i]) != NONE32) own = 0;
        if (i < v_lo) own = 0; // first created by a ghost tet of the rank below (rin_set_ghost_tets)
        own_flag[i] = own;
    }
}

// single-block exclusive scan of 0/1 flags -> own_idx (NONE32 where the flag is 0); total in *n_own
__global__ void __launch_bounds__(1024) own_scan_kernel(const uint32_t* __restrict__ flag, uint32_t n,
    uint32_t* __restrict__ own_idx, unsigned* __restrict__ n_own)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t c = (i < n) ? flag[i] : 0;
        uint32_t x = c;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t t = s_w[lane], y = t;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t z = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= o) y += z;
            }
            s_w[lane] = y - t;
        }
        __syncthreads();
        const uint32_t r = s_run, wv = s_w[warp];
        if (i < n) own_idx[i] = c ? (r + wv + x - c) : NONE32;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = r + wv + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_own = s_run;
}

// gid[i] = global id of local vertex i: offset + own index, or the owner's id for foreign vertices
__global__ void __launch_bounds__(256) global_ids_kernel(const uint4* __restrict__ v_key,
    const uint32_t* __restrict__ own_idx, uint32_t n, uint32_t offset, const uint4* __restrict__ fkeys,
    const uint32_t* __restrict__ fgids, const uint32_t* __restrict__ table, uint32_t mask,
    uint32_t* __restrict__ gid, unsigned* __restrict__ n_unresolved, uint32_t v_lo)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t o = own_idx[i];
        if (o != NONE32) {
            gid[i] = offset + o;
            continue;
        }
        const uint32_t f = mask ? foreign_lookup(fkeys, table, mask, v_key[i]) : NONE32;
        if (f == NONE32) {
            if (i >= v_lo) atomicAdd(n_unresolved, 1u); // a ghost's vertex off the shared plane: unused
            gid[i] = NONE32;
        } else
            gid[i] = fgids[f];
    }
}

__global__ void __launch_bounds__(256) apply_gids_kernel(uint32_t* __restrict__ f_verts, uint32_t n,
    const uint32_t* __restrict__ gid)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        f_verts[i] = gid[f_verts[i]];
}

// ghost runs: a kept face must not use a vertex that has no global id
__global__ void __launch_bounds__(256) check_gids_kernel(const uint32_t* __restrict__ f_verts, uint32_t n,
    const uint32_t* __restrict__ gid, unsigned* __restrict__ n_bad)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (gid[f_verts[i]] == NONE32) atomicAdd(n_bad, 1u);
}

// keep the owned vertices (order preserved)
__global__ void __launch_bounds__(256) compact_own_verts_kernel(const uint32_t* __restrict__ own_idx, uint32_t n,
    const uint32_t* __restrict__ v_tet, const uint8_t* __restrict__ v_local, const uint8_t* __restrict__ v_size,
    const uint4* __restrict__ v_simplex, const uint4* __restrict__ v_funcs, const double* __restrict__ v_xyz,
    const uint4* __restrict__ v_key, uint32_t* __restrict__ o_tet, uint8_t* __restrict__ o_local,
    uint8_t* __restrict__ o_size, uint4* __restrict__ o_simplex, uint4* __restrict__ o_funcs,
    double* __restrict__ o_xyz, uint4* __restrict__ o_key, const unsigned* __restrict__ n_dev = nullptr)
{
    if (n_dev) n = *n_dev;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t d = own_idx[i];
        if (d == NONE32) continue;
        o_tet[d] = v_tet[i];
        o_local[d] = v_local[i];
        o_size[d] = v_size[i];
        o_simplex[d] = v_simplex[i];
        o_funcs[d] = v_funcs[i];
        o_xyz[3 * (size_t)d] = v_xyz[3 * (size_t)i];
        o_xyz[3 * (size_t)d + 1] = v_xyz[3 * (size_t)i + 1];
        o_xyz[3 * (size_t)d + 2] = v_xyz[3 * (size_t)i + 2];
        o_key[d] = v_key[i];
    }
}

} // namespace rin
