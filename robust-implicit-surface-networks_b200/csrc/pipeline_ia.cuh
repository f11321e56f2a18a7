// The implicit-arrangement pass as nine launches with no host synchronisation in between:
//
//   eval_*_kernel            K1  function values (SoA) + per-vertex sign masks
//   filter_classify_kernel   K2  active-function filter + table dispatch, tile-local ordered compaction
//   general_ia_small_kernel  K4  general per-tet arrangement, <= 4 functions (kernels_ia.cuh)
//   general_ia_mid_kernel    K4  more functions / overflows (kernels_ia.cuh)
//   scan_tiles5_kernel           exclusive scan of the per-tile totals (one block of 32 warps)
//   emit_kernel              K5  ordered active list, canonical vertex keys
//   insert_kernel            K6  hash-min insertion of the shareable candidates
//   rank_verts_kernel        K6+K7  first-occurrence ranking, IsoVert records, coordinates
//   faces_kernel             K5  PolygonFace arrays with final vertex ids
//
// Reference stages replaced: src/implicit_arrangement.cpp:61-116 (signs, filter), :244-306 (dispatch),
// src/extract_mesh.cpp:10-265 (extract_iso_mesh), :1446-1538 (compute_iso_vert_xyz).
// Buffers are sized from the previous pass (or from a sizing pass on first use); every kernel checks the
// capacities on the device and the single read-back at the end tells the host whether to grow and repeat.
#pragma once
#include "kernels_ia.cuh"

namespace rin {

// ---------------------------------------------------------------------------------------------
// K0': coordinate tables of a generated grid: xs[i] = (i / (N-1)) * (max - min) + min, the
// expression of src/io.cpp:104-113 (the device points are built from the same expression).
// ---------------------------------------------------------------------------------------------
__global__ void grid_axes_kernel(uint32_t N, double3 bmin, double3 bmax, double* __restrict__ axes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double d = double(N - 1);
    axes[i] = (double(i) / d) * (bmax.x - bmin.x) + bmin.x;
    axes[N + i] = (double(i) / d) * (bmax.y - bmin.y) + bmin.y;
    axes[2 * N + i] = (double(i) / d) * (bmax.z - bmin.z) + bmin.z;
}

// vertex ids of tet t of generate_tet_mesh(R) (src/io.cpp:122-147); d5 / dR divide by 5 / R
__device__ __forceinline__ uint4 grid_tet(uint32_t R, const FastDiv dR, uint32_t t)
{
    const uint32_t N = R + 1;
    const uint32_t cube = t / 5, s = t - 5 * cube;
    const uint32_t ij = fd_div(cube, dR), k = cube - ij * R, i = fd_div(ij, dR), j = ij - i * R;
    const uint8_t* tab = ((i + j + k) & 1) ? c_grid_odd[s] : c_grid_even[s];
    uint32_t v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t q = tab[c];
        const uint32_t di = ((q & 3) == 1 || (q & 3) == 2), dj = ((q & 3) >= 2), dk = (q >> 2);
        v[c] = (i + di) * N * N + (j + dj) * N + (k + dk);
    }
    return make_uint4(v[0], v[1], v[2], v[3]);
}

// ---------------------------------------------------------------------------------------------
// K1: evaluation + sign masks, EV_VPT consecutive vertices per thread (128-bit stores): the function kind is
// branched on once per function (warp-uniform), one descriptor fetch serves all the thread's vertices.
// GRID: coordinates come from the three axis tables, nothing else is read.
// PACK (F <= 16): only the packed word P | N << 16 is written, else the (P, N) pairs per 32 functions.
// VS = row stride of vals / vmask (V rounded up so that every row starts on a 128-byte boundary).
// The arithmetic is that of eval_func (kernels_ia.cuh), operation for operation.
// ---------------------------------------------------------------------------------------------
constexpr int EV_VPT = 4;

template <int NVX = EV_VPT>
__device__ __forceinline__ void eval_func_n(const rin_func_desc& f, const double* x, const double* y,
    const double* z, double* v)
{
    switch (f.type) {
    case RIN_FN_PLANE: {
        const double p0 = f.p[0], p1 = f.p[1], p2 = f.p[2], n0 = f.p[3], n1 = f.p[4], n2 = f.p[5];
#pragma unroll
        for (int j = 0; j < NVX; ++j) {
            const double dx = x[j] - p0, dy = y[j] - p1, dz = z[j] - p2;
            v[j] = (n0 * dx + n1 * dy) + n2 * dz;
        }
        break;
    }
    case RIN_FN_SPHERE: {
        const double p0 = f.p[0], p1 = f.p[1], p2 = f.p[2], r = f.p[3];
        if (f.p[4] != 0.0) { // squared
            const double r2 = r * r;
#pragma unroll
            for (int j = 0; j < NVX; ++j) {
                const double dx = x[j] - p0, dy = y[j] - p1, dz = z[j] - p2;
                v[j] = r2 - ((dx * dx + dy * dy) + dz * dz);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NVX; ++j) {
                const double dx = x[j] - p0, dy = y[j] - p1, dz = z[j] - p2;
                v[j] = r - sqrt((dx * dx + dy * dy) + dz * dz);
            }
        }
        break;
    }
    case RIN_FN_CYLINDER:
    case RIN_FN_TORUS: {
        const double p0 = f.p[0], p1 = f.p[1], p2 = f.p[2], a0 = f.p[3], a1 = f.p[4], a2 = f.p[5], r = f.p[6];
        const bool torus = f.type == RIN_FN_TORUS;
        const double r2 = f.p[7];
#pragma unroll
        for (int j = 0; j < NVX; ++j) {
            const double dx = x[j] - p0, dy = y[j] - p1, dz = z[j] - p2;
            const double t = (a0 * dx + a1 * dy) + a2 * dz;
            const double px = dx - t * a0, py = dy - t * a1, pz = dz - t * a2;
            const double d = sqrt((px * px + py * py) + pz * pz);
            if (torus) {
                const double rho = d - r;
                v[j] = r2 - sqrt(rho * rho + t * t);
            } else
                v[j] = r - d;
        }
        break;
    }
    default:
#pragma unroll
        for (int j = 0; j < NVX; ++j) v[j] = 0.0;
        break;
    }
    if (f.flip) {
#pragma unroll
        for (int j = 0; j < NVX; ++j) v[j] = -v[j];
    }
}

template <bool GRID>
__global__ void __launch_bounds__(256) eval_kernel(const double* __restrict__ pts, const double* __restrict__ axes,
    uint32_t N, const FastDiv dN, uint32_t v_first, uint32_t v_count, uint32_t VS,
    const rin_func_desc* __restrict__ funcs, uint32_t F, int negate, double* __restrict__ vals,
    uint2* __restrict__ vmask, uint32_t* __restrict__ vmask16, unsigned long long* __restrict__ n_zero)
{
    extern __shared__ rin_func_desc s_funcs[];
    for (uint32_t i = threadIdx.x; i < F * (sizeof(rin_func_desc) / 8); i += blockDim.x)
        reinterpret_cast<double*>(s_funcs)[i] = reinterpret_cast<const double*>(funcs)[i];
    __syncthreads();
    unsigned zeros = 0;
    // A thread owns EV_VPT = 4 CONSECUTIVE vertices starting at a multiple of 4: its values of one function are 32
    // contiguous bytes of the row (two 128-bit stores, a warp writes 1 KB contiguously), its packed masks one 128-bit
    // store; on a generated grid the four vertices share one index decomposition unless the run crosses a grid line.
    const uint32_t v_end = v_first + v_count;
    const uint32_t g_first = v_first / EV_VPT, g_end = (v_end + EV_VPT - 1) / EV_VPT;
    for (uint32_t g = g_first + blockIdx.x * blockDim.x + threadIdx.x; g < g_end; g += gridDim.x * blockDim.x) {
        const uint32_t vb = g * EV_VPT;
        double x[EV_VPT], y[EV_VPT], z[EV_VPT];
        bool ok[EV_VPT];
        bool all = true;
#pragma unroll
        for (int j = 0; j < EV_VPT; ++j) {
            ok[j] = vb + j >= v_first && vb + j < v_end;
            all &= ok[j];
        }
        if (GRID) {
            const uint32_t ij = fd_div(vb, dN), k = vb - ij * N, ii = fd_div(ij, dN), jj = ij - ii * N;
            if (k + EV_VPT <= N) {
                const double xs = __ldg(&axes[ii]), ys = __ldg(&axes[N + jj]);
#pragma unroll
                for (int j = 0; j < EV_VPT; ++j) {
                    x[j] = xs;
                    y[j] = ys;
                    z[j] = __ldg(&axes[2 * N + k + j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < EV_VPT; ++j) {
                    const uint32_t v = ok[j] ? vb + j : v_first;
                    const uint32_t ij2 = fd_div(v, dN), k2 = v - ij2 * N, i2 = fd_div(ij2, dN), j2 = ij2 - i2 * N;
                    x[j] = __ldg(&axes[i2]);
                    y[j] = __ldg(&axes[N + j2]);
                    z[j] = __ldg(&axes[2 * N + k2]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < EV_VPT; ++j) {
                const size_t v = ok[j] ? vb + j : v_first;
                x[j] = __ldg(&pts[3 * v]);
                y[j] = __ldg(&pts[3 * v + 1]);
                z[j] = __ldg(&pts[3 * v + 2]);
            }
        }
        for (uint32_t w = 0; w * 32 < F; ++w) {
            uint32_t P[EV_VPT], Nn[EV_VPT];
#pragma unroll
            for (int j = 0; j < EV_VPT; ++j) P[j] = Nn[j] = 0;
            const uint32_t fe = min(F, w * 32 + 32);
            for (uint32_t f = w * 32; f < fe; ++f) {
                double val[EV_VPT];
                eval_func_n(s_funcs[f], x, y, z, val);
                double* __restrict__ row = vals + (size_t)f * VS + vb;
                const uint32_t bit = 1u << (f & 31);
#pragma unroll
                for (int j = 0; j < EV_VPT; ++j) {
                    if (negate) val[j] = val[j] * -1; // csg(): funcVals * -1 (src/csg.cpp:37)
                    if (val[j] > 0) P[j] |= bit;
                    if (val[j] < 0) Nn[j] |= bit;
                }
                if (all) {
                    reinterpret_cast<double2*>(row)[0] = make_double2(val[0], val[1]);
                    reinterpret_cast<double2*>(row)[1] = make_double2(val[2], val[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < EV_VPT; ++j)
                        if (ok[j]) row[j] = val[j];
                }
            }
            if (vmask16) {
                if (all)
                    *reinterpret_cast<uint4*>(vmask16 + vb) = make_uint4(P[0] | (Nn[0] << 16), P[1] | (Nn[1] << 16),
                        P[2] | (Nn[2] << 16), P[3] | (Nn[3] << 16));
                else {
#pragma unroll
                    for (int j = 0; j < EV_VPT; ++j)
                        if (ok[j]) vmask16[vb + j] = P[j] | (Nn[j] << 16);
                }
            } else {
#pragma unroll
                for (int j = 0; j < EV_VPT; ++j)
                    if (ok[j]) vmask[(size_t)w * VS + vb + j] = make_uint2(P[j], Nn[j]);
            }
#pragma unroll
            for (int j = 0; j < EV_VPT; ++j)
                if (ok[j]) zeros += (fe - w * 32) - __popc(P[j] | Nn[j]);
        }
    }
    // num_degenerate_vertex counts (vertex, function) pairs with value 0 (src/implicit_arrangement.cpp:69-73)
    for (int o = 16; o; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(n_zero, (unsigned long long)zeros);
}

// ---------------------------------------------------------------------------------------------
// K1 (material interface, up to MI_EVAL_MAXF materials): evaluation fused with the "highest func" loop
// (src/material_interface.cpp:59-92).  The F values of a vertex stay in registers, so the maximal-material mask,
// the "some two materials are exactly equal" flag (tie gate) and the tie count cost no second pass over the
// 8F bytes per vertex that highest_material_kernel re-reads.  Same arithmetic as eval_kernel / eval_func.
// ---------------------------------------------------------------------------------------------
constexpr int MI_EVAL_MAXF = 8;
constexpr int MI_EVAL_VPT = 4;
__host__ __device__ constexpr size_t mi_eval_smem(uint32_t F)
{
    // function descriptors (rounded up to 16 bytes) + the block's values [material][vertex slot][thread]
    return ((F * sizeof(rin_func_desc) + 15) & ~(size_t)15) + (size_t)F * MI_EVAL_VPT * 256 * sizeof(double);
}
template <bool GRID>
__global__ void __launch_bounds__(256) eval_mi_kernel(const double* __restrict__ pts, const double* __restrict__ axes,
    uint32_t N, const FastDiv dN, uint32_t v_first, uint32_t v_count, uint32_t VS,
    const rin_func_desc* __restrict__ funcs, uint32_t F, int negate, double* __restrict__ vals,
    uint2* __restrict__ vmask, unsigned long long* __restrict__ n_tied)
{
    extern __shared__ __align__(16) rin_func_desc s_funcs[];
    double* s_val = reinterpret_cast<double*>(
        reinterpret_cast<uint8_t*>(s_funcs) + ((F * sizeof(rin_func_desc) + 15) & ~(size_t)15));
    for (uint32_t i = threadIdx.x; i < F * (sizeof(rin_func_desc) / 8); i += blockDim.x)
        reinterpret_cast<double*>(s_funcs)[i] = reinterpret_cast<const double*>(funcs)[i];
    __syncthreads();
    unsigned tied = 0;
    const uint32_t step = gridDim.x * blockDim.x * MI_EVAL_VPT;
    for (uint32_t i0 = blockIdx.x * blockDim.x * MI_EVAL_VPT + threadIdx.x; i0 < v_count; i0 += step) {
        double x[MI_EVAL_VPT], y[MI_EVAL_VPT], z[MI_EVAL_VPT];
        uint32_t v[MI_EVAL_VPT];
        bool ok[MI_EVAL_VPT];
#pragma unroll
        for (int j = 0; j < MI_EVAL_VPT; ++j) {
            const uint32_t idx = i0 + j * blockDim.x;
            ok[j] = idx < v_count;
            v[j] = v_first + (ok[j] ? idx : 0u);
            if (GRID) {
                const uint32_t ij = fd_div(v[j], dN), k = v[j] - ij * N, ii = fd_div(ij, dN), jj = ij - ii * N;
                x[j] = __ldg(&axes[ii]);
                y[j] = __ldg(&axes[N + jj]);
                z[j] = __ldg(&axes[2 * N + k]);
            } else {
                x[j] = __ldg(&pts[3 * (size_t)v[j]]);
                y[j] = __ldg(&pts[3 * (size_t)v[j] + 1]);
                z[j] = __ldg(&pts[3 * (size_t)v[j] + 2]);
            }
        }
        // running maximum and the set of materials attaining it (first index on ties is bit order), :66-76;
        // the values are parked in the thread's own shared-memory column for the pairwise test below
        double mx[MI_EVAL_VPT];
        uint32_t H[MI_EVAL_VPT];
        for (uint32_t f = 0; f < F; ++f) {
            double tmp[MI_EVAL_VPT];
            eval_func_n<MI_EVAL_VPT>(s_funcs[f], x, y, z, tmp);
            double* __restrict__ row = vals + (size_t)f * VS;
#pragma unroll
            for (int j = 0; j < MI_EVAL_VPT; ++j) {
                if (negate) tmp[j] = tmp[j] * -1;
                if (ok[j]) row[v[j]] = tmp[j];
                s_val[(f * MI_EVAL_VPT + j) * 256 + threadIdx.x] = tmp[j];
                if (f == 0 || tmp[j] > mx[j] || mx[j] != mx[j]) {
                    mx[j] = tmp[j];
                    H[j] = 1u << f;
                } else if (tmp[j] == mx[j])
                    H[j] |= 1u << f;
            }
        }
#pragma unroll
        for (int j = 0; j < MI_EVAL_VPT; ++j) {
            if (!ok[j]) continue;
            // "some two materials are exactly equal here" (tie gate of classify_mi_kernel)
            // all pairs, fully unrolled; the slots of absent materials hold NaN, which equals nothing
            double a[MI_EVAL_MAXF];
#pragma unroll
            for (int f = 0; f < MI_EVAL_MAXF; ++f)
                a[f] = ((uint32_t)f < F) ? s_val[(f * MI_EVAL_VPT + j) * 256 + threadIdx.x]
                                         : __longlong_as_double(0x7ff8000000000000ll);
            bool any_equal = false;
#pragma unroll
            for (int f = 0; f < MI_EVAL_MAXF; ++f)
#pragma unroll
                for (int g = f + 1; g < MI_EVAL_MAXF; ++g) any_equal |= (a[f] == a[g]);
            if (mx[j] != mx[j]) H[j] = 0; // every value NaN: nothing compares equal to the maximum
            vmask[v[j]] = make_uint2(H[j], any_equal ? 1u : 0u);
            tied += (__popc(H[j]) > 1);
        }
    }
    for (int o = 16; o; o >>= 1) tied += __shfl_xor_sync(0xffffffffu, tied, o);
    if ((threadIdx.x & 31) == 0 && tied) atomicAdd(n_tied, (unsigned long long)tied);
}

// ---------------------------------------------------------------------------------------------
// K2: filter + dispatch.  One streaming pass, no inter-block dependency: every block compacts the
// active tets of its tile into tile-local slots (tet id, function masks, record reference) in tet
// order and writes the tile's totals; tets that need the general algorithm are appended to the work
// lists with their slot.
//   GRID: a thread owns a whole cube of generate_tet_mesh (five tets) and loads its eight corner masks
//         once, neighbouring threads read neighbouring words; the tet index stream is never read.
//   else: the 16-byte index records are streamed and the four masks gathered.
// Filter rule: function j is active iff not positive at all four corners and not negative at all four
// (pos < 4 && neg < 4, src/implicit_arrangement.cpp:106).  Dispatch rules (:276-284): tables serve 1
// function, and 2 functions when the secondary lookup is on; a zero at a corner, coincident crossing
// points or a missing entry take the general kernel.
// ---------------------------------------------------------------------------------------------
struct FilterArgs
{
    const uint4* tets; // explicit index records (null for GRID)
    uint32_t R;        // GRID: resolution
    FastDiv dR;        // division by R
    uint32_t t_first, t_count;
    uint32_t c_first, n_units; // GRID: first cube / number of cubes; else n_units = t_count
    const uint2* vmask;
    const uint32_t* vmask16;
    uint32_t VS;
    uint32_t last_mask;
    const double* vals;
    const uint16_t* lut1;
    const uint16_t* lut2;
    const uint32_t* blob32;
    uint32_t* arena32;
    int use_lookup, use_secondary;
    uint32_t* tl_tet;
    uint32_t* tl_mask;
    uint32_t* tl_ref;
    size_t tl_stride;
    TileTot* tile_tot;
    uint32_t* small_list;
    uint32_t* big_list;
    uint32_t list_cap;
    FilterCounters* fc;
    GeneralCounters* gc;
    unsigned* n_exact;
    unsigned* overflow;
};

__device__ __forceinline__ unsigned agg_inc(unsigned* ctr)
{
    const unsigned am = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(am) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned)__popc(am));
    base = __shfl_sync(am, base, leader);
    return base + __popc(am & ((1u << lane) - 1u));
}

template <int W, bool PACK>
__device__ __forceinline__ uint32_t classify_ia_tet(const FilterArgs& A, const uint4 tv, const uint32_t* m, int k,
    unsigned& exact)
{
    if (!A.use_lookup || k > 2 || (k == 2 && !A.use_secondary)) return REF_GENERAL;
    const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
    const int f0 = (W == 1) ? (__ffs(m[0]) - 1) : nth_set_bit(m, W, 0);
    const int f1 = (k == 2) ? nth_set_bit(m, W, 1) : f0;
    int s0 = 0, s1 = 0;
    bool nonzero = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t P0, Z0, P1, Z1;
        if (PACK) {
            const uint32_t x = __ldg(&A.vmask16[vv[c]]);
            P0 = (x >> f0) & 1;
            Z0 = ((x | (x >> 16)) >> f0) & 1;
            P1 = (x >> f1) & 1;
            Z1 = ((x | (x >> 16)) >> f1) & 1;
        } else {
            const uint2 a0 = __ldg(&A.vmask[(size_t)(f0 >> 5) * A.VS + vv[c]]);
            P0 = (a0.x >> (f0 & 31)) & 1;
            Z0 = ((a0.x | a0.y) >> (f0 & 31)) & 1;
            const uint2 a1 = (k == 2) ? __ldg(&A.vmask[(size_t)(f1 >> 5) * A.VS + vv[c]]) : a0;
            P1 = (a1.x >> (f1 & 31)) & 1;
            Z1 = ((a1.x | a1.y) >> (f1 & 31)) & 1;
        }
        s0 |= P0 << c;
        s1 |= P1 << c;
        nonzero &= (Z0 & Z1) != 0;
    }
    if (!nonzero) return REF_GENERAL;
    if (k == 1) {
        const uint16_t off = __ldg(&A.lut1[s0]);
        return off == LUT_MISS ? REF_GENERAL : (uint32_t)off;
    }
    double p0[4], p1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        p0[c] = __ldg(&A.vals[(size_t)f0 * A.VS + vv[c]]);
        p1[c] = __ldg(&A.vals[(size_t)f1 * A.VS + vv[c]]);
    }
    const int key = ia2_key(p0, p1, &exact);
    if (key < 0) return REF_GENERAL;
    const uint16_t off = __ldg(&A.lut2[key]);
    return off == LUT_MISS ? REF_GENERAL : (uint32_t)off;
}

__host__ __device__ constexpr int filter_rounds(int W, bool grid)
{
    return grid ? (W == 1 ? 4 : (W == 2 ? 2 : 1)) : (W == 1 ? 8 : (W == 2 ? 4 : 2));
}
__host__ __device__ constexpr int filter_tile_slots(int W, bool grid)
{
    return filter_rounds(W, grid) * 256 * (grid ? 5 : 1);
}

template <int W, bool PACK, bool GRID>
__global__ void __launch_bounds__(256) filter_classify_kernel(const FilterArgs A)
{
    constexpr int ROUNDS = filter_rounds(W, GRID);
    constexpr int PER = GRID ? 5 : 1;
    constexpr int UNITS = ROUNDS * 256;
    constexpr int TS = UNITS * PER;
    constexpr int NE = ROUNDS * 8; // (round, warp) entries
    __shared__ unsigned s_cnt[NE];
    __shared__ unsigned s_tot[8];
    const unsigned tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 8) s_tot[threadIdx.x] = 0;
    if (tile == 0 && threadIdx.x == 0) {
        // the arena of general records starts with an empty record (header + trailing word)
        A.arena32[0] = 0;
        A.arena32[1] = 0;
        A.gc->arena_top = 8;
    }
    const uint32_t t_end = A.t_first + A.t_count;
    const uint32_t N = A.R + 1;

    // ---- masks of every tet of the tile (all loads of the tile are independent)
    uint32_t m[ROUNDS][PER][W];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t u = tile * UNITS + r * 256 + threadIdx.x;
        const bool in = u < A.n_units;
        if constexpr (GRID) {
            const uint32_t cube = A.c_first + (in ? u : 0u);
            const uint32_t ij = fd_div(cube, A.dR), k = cube - ij * A.R, i = fd_div(ij, A.dR), j = ij - i * A.R;
            const uint32_t base = (i * N + j) * N + k;
            const bool odd = (i + j + k) & 1;
            // corner q of the cube, numbered like v0..v7 at src/io.cpp:126-133
            uint32_t cv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                cv[q] = base + (((q & 3) == 1 || (q & 3) == 2) ? N * N : 0u) + (((q & 3) >= 2) ? N : 0u) + (q >> 2);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t mm[5];
                if (PACK) {
                    uint32_t c8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) c8[q] = __ldg(&A.vmask16[cv[q]]);
#define RIN_TM(a, b, c, d) (c8[a] & c8[b] & c8[c] & c8[d])
                    uint32_t x[5];
                    if (!odd) {
                        x[0] = RIN_TM(4, 6, 1, 3);
                        x[1] = RIN_TM(6, 3, 4, 7);
                        x[2] = RIN_TM(1, 3, 0, 4);
                        x[3] = RIN_TM(3, 1, 2, 6);
                        x[4] = RIN_TM(4, 1, 6, 5);
                    } else {
                        x[0] = RIN_TM(7, 0, 2, 5);
                        x[1] = RIN_TM(2, 3, 0, 7);
                        x[2] = RIN_TM(5, 7, 0, 4);
                        x[3] = RIN_TM(7, 2, 6, 5);
                        x[4] = RIN_TM(0, 1, 2, 5);
                    }
#undef RIN_TM
#pragma unroll
                    for (int s = 0; s < 5; ++s) mm[s] = ~((x[s] & 0xffffu) | (x[s] >> 16));
                } else {
                    uint2 c8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) c8[q] = __ldg(&A.vmask[(size_t)w * A.VS + cv[q]]);
#define RIN_TM(a, b, c, d) \
    ~((c8[a].x & c8[b].x & c8[c].x & c8[d].x) | (c8[a].y & c8[b].y & c8[c].y & c8[d].y))
                    if (!odd) {
                        mm[0] = RIN_TM(4, 6, 1, 3);
                        mm[1] = RIN_TM(6, 3, 4, 7);
                        mm[2] = RIN_TM(1, 3, 0, 4);
                        mm[3] = RIN_TM(3, 1, 2, 6);
                        mm[4] = RIN_TM(4, 1, 6, 5);
                    } else {
                        mm[0] = RIN_TM(7, 0, 2, 5);
                        mm[1] = RIN_TM(2, 3, 0, 7);
                        mm[2] = RIN_TM(5, 7, 0, 4);
                        mm[3] = RIN_TM(7, 2, 6, 5);
                        mm[4] = RIN_TM(0, 1, 2, 5);
                    }
#undef RIN_TM
                }
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    const uint32_t t = 5 * cube + s;
                    uint32_t x = mm[s];
                    if (w == W - 1) x &= A.last_mask;
                    if (!in || t < A.t_first || t >= t_end) x = 0;
                    m[r][s][w] = x;
                }
            }
        } else {
            const uint4 tv = __ldg(&A.tets[A.t_first + (in ? u : 0u)]);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t x;
                if (PACK) {
                    const uint32_t g = __ldg(&A.vmask16[tv.x]) & __ldg(&A.vmask16[tv.y]) & __ldg(&A.vmask16[tv.z]) &
                                       __ldg(&A.vmask16[tv.w]);
                    x = ~((g & 0xffffu) | (g >> 16));
                } else {
                    const uint2 g0 = __ldg(&A.vmask[(size_t)w * A.VS + tv.x]);
                    const uint2 g1 = __ldg(&A.vmask[(size_t)w * A.VS + tv.y]);
                    const uint2 g2 = __ldg(&A.vmask[(size_t)w * A.VS + tv.z]);
                    const uint2 g3 = __ldg(&A.vmask[(size_t)w * A.VS + tv.w]);
                    x = ~((g0.x & g1.x & g2.x & g3.x) | (g0.y & g1.y & g2.y & g3.y));
                }
                if (w == W - 1) x &= A.last_mask;
                if (!in) x = 0;
                m[r][0][w] = x;
            }
        }
    }

    // ---- ordered ranks: lane prefix inside (round, warp), then the prefix over the (round, warp) sequence
    unsigned pr[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        unsigned cnt = 0;
#pragma unroll
        for (int s = 0; s < PER; ++s) {
            uint32_t any = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) any |= m[r][s][w];
            cnt += any != 0;
        }
        unsigned x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        pr[r] = x - cnt;
        if (lane == 31) s_cnt[r * 8 + warp] = x;
    }
    __syncthreads();
    unsigned my_off[ROUNDS];
    unsigned run = 0;
#pragma unroll
    for (int b = 0; b < (NE + 31) / 32; ++b) {
        const int e = b * 32 + lane;
        const unsigned c = e < NE ? s_cnt[e] : 0u;
        unsigned x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        const unsigned excl = run + x - c;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int idx = r * 8 + warp;
            if (idx / 32 == b) my_off[r] = __shfl_sync(0xffffffffu, excl, idx % 32);
        }
        run += __shfl_sync(0xffffffffu, x, 31);
    }

    // ---- tile-local slots (tet id, masks) in tet order; most warps have nothing to write
    const size_t tbase = (size_t)tile * TS;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if (s_cnt[r * 8 + warp] == 0) continue;
        unsigned rank = 0;
#pragma unroll
        for (int s = 0; s < PER; ++s) {
            uint32_t any = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) any |= m[r][s][w];
            if (any) {
                const uint32_t u = tile * UNITS + r * 256 + threadIdx.x;
                const size_t pos = tbase + my_off[r] + pr[r] + rank++;
                A.tl_tet[pos] = GRID ? 5 * (A.c_first + u) + s : A.t_first + u;
#pragma unroll
                for (int w = 0; w < W; ++w) A.tl_mask[(size_t)w * A.tl_stride + pos] = m[r][s][w];
            }
        }
    }
    __syncthreads(); // the block's own global writes are visible to the block from here

    // ---- dispatch, dense over the tile's active slots (all lanes busy, one copy of the code)
    unsigned k1 = 0, k2 = 0, km = 0, kf = 0, nc = 0, nfa = 0, nfv = 0, exact = 0;
    for (uint32_t i = threadIdx.x; i < run; i += 256) {
        const size_t pos = tbase + i;
        const uint32_t t = __ldcg(&A.tl_tet[pos]);
        uint32_t mw[W];
        int k = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            mw[w] = __ldcg(&A.tl_mask[(size_t)w * A.tl_stride + pos]);
            k += __popc(mw[w]);
        }
        k1 += (k == 1);
        k2 += (k == 2);
        km += (k > 2);
        kf += k;
        const uint4 tv = GRID ? grid_tet(A.R, A.dR, t) : __ldg(&A.tets[t]);
        const uint32_t ref = classify_ia_tet<W, PACK>(A, tv, mw, k, exact);
        A.tl_ref[pos] = ref;
        if (ref & REF_GENERAL) {
            if (k <= IACapsSmall::MAXK) {
                const unsigned q = agg_inc(&A.gc->n_small);
                if (q < A.list_cap)
                    A.small_list[q] = (uint32_t)pos;
                else
                    atomicOr(A.overflow, OVF_LIST);
            } else {
                const unsigned q = agg_inc(&A.gc->n_big);
                if (q < A.list_cap)
                    A.big_list[q] = (uint32_t)pos;
                else
                    atomicOr(A.overflow, OVF_LIST);
            }
        } else {
            const uint32_t h = __ldg(&A.blob32[ref]);
            nc += h & 255;
            nfa += (h >> 8) & 255;
            nfv += h >> 16;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        k1 += __shfl_xor_sync(0xffffffffu, k1, o);
        k2 += __shfl_xor_sync(0xffffffffu, k2, o);
        km += __shfl_xor_sync(0xffffffffu, km, o);
        kf += __shfl_xor_sync(0xffffffffu, kf, o);
        nc += __shfl_xor_sync(0xffffffffu, nc, o);
        nfa += __shfl_xor_sync(0xffffffffu, nfa, o);
        nfv += __shfl_xor_sync(0xffffffffu, nfv, o);
        exact += __shfl_xor_sync(0xffffffffu, exact, o);
    }
    if (lane == 0) {
        if (k1) atomicAdd(&s_tot[0], k1);
        if (k2) atomicAdd(&s_tot[1], k2);
        if (km) atomicAdd(&s_tot[2], km);
        if (kf) atomicAdd(&s_tot[3], kf);
        if (nc) atomicAdd(&s_tot[4], nc);
        if (nfa) atomicAdd(&s_tot[5], nfa);
        if (nfv) atomicAdd(&s_tot[6], nfv);
        if (exact) atomicAdd(&s_tot[7], exact);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint4* tt = reinterpret_cast<uint4*>(A.tile_tot + tile);
        tt[0] = make_uint4(run, s_tot[3], s_tot[4], s_tot[5]);
        tt[1] = make_uint4(s_tot[6], 0, 0, 0);
        if (s_tot[0]) atomicAdd(&A.fc->n_k1, s_tot[0]);
        if (s_tot[1]) atomicAdd(&A.fc->n_k2, s_tot[1]);
        if (s_tot[2]) atomicAdd(&A.fc->n_kmore, s_tot[2]);
        if (s_tot[7]) atomicAdd(A.n_exact, s_tot[7]);
    }
}

// ---------------------------------------------------------------------------------------------
// Canonical description of one iso-vertex of a tet (the classification of src/extract_mesh.cpp:93-225):
// by the number of simplex-boundary planes among its three planes it lies
//   on a tet vertex (3): key = grid vertex                       :197-225
//   on a tet edge   (2): key = (vmin, vmax, fn)                   :113-151
//   on a tet face   (1): key = (v0 <= v1 <= v2, fn_a, fn_b)       :153-183
//   inside          (0): never shared                             :185-195
// key  = (v0, v1, v2, fa | fb << 16), unused slots 0xffffffff / 0xffff; key.x = NONE32 marks "never shared"
// size = number of simplex vertices; f[3] = function ids in plane order (0xffff = unused)
// ---------------------------------------------------------------------------------------------
struct IsoVertInfo
{
    uint4 key;
    uint32_t f[3];
    int size, local;
};

template <int W>
__device__ __forceinline__ uint32_t func_of_plane(int pl, const uint32_t* fl, const uint32_t* m)
{
    const int j = pl - 4;
    return j < 4 ? fl[j] : (uint32_t)nth_set_bit(m, W, j);
}

template <int W>
__device__ __forceinline__ void first_funcs(const uint32_t* m, uint32_t* fl)
{
    fl[0] = fl[1] = fl[2] = fl[3] = 0xffffu;
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

template <int W>
__device__ __forceinline__ IsoVertInfo iso_vert_info(uint32_t e, const uint32_t* tv, const uint32_t* fl,
    const uint32_t* m)
{
    IsoVertInfo r;
    r.local = e & 255;
    const int p0 = (e >> 8) & 255, p1 = (e >> 16) & 255, p2 = e >> 24;
    // planes ascend: boundary planes (< 4) come first
    const int nb = (p0 < 4) + (p1 < 4) + (p2 < 4);
    r.f[0] = r.f[1] = r.f[2] = 0xffffu;
    if (nb == 0) {
        r.f[0] = func_of_plane<W>(p0, fl, m);
        r.f[1] = func_of_plane<W>(p1, fl, m);
        r.f[2] = func_of_plane<W>(p2, fl, m);
        r.key = make_uint4(NONE32, tv[1], tv[2], tv[3]);
        r.size = 4;
        return r;
    }
    uint32_t c0, c1 = NONE32, c2 = NONE32;
    if (nb == 3) { // on the tet corner not listed
        c0 = tv[6 - p0 - p1 - p2];
        r.size = 1;
    } else if (nb == 2) { // on the tet edge between the two corners not listed
        r.f[0] = func_of_plane<W>(p2, fl, m);
        const unsigned rest = 0xfu & ~((1u << p0) | (1u << p1));
        const int a0 = __ffs(rest) - 1, a1 = 31 - __clz(rest);
        c0 = min(tv[a0], tv[a1]);
        c1 = max(tv[a0], tv[a1]);
        r.size = 2;
    } else { // on the tet face opposite corner p0
        r.f[0] = func_of_plane<W>(p1, fl, m);
        r.f[1] = func_of_plane<W>(p2, fl, m);
        const uint32_t x = tv[p0 == 0 ? 1 : 0], y = tv[p0 <= 1 ? 2 : 1], z = tv[p0 == 3 ? 2 : 3];
        const uint32_t lo = min(x, min(y, z)), hi = max(x, max(y, z));
        c0 = lo;
        c1 = x ^ y ^ z ^ lo ^ hi; // the middle one
        c2 = hi;
        r.size = 3;
    }
    r.key = make_uint4(c0, c1, c2, r.f[0] | (r.f[1] << 16)); // function ids in plane order (:138,:167-168)
    return r;
}

// ---------------------------------------------------------------------------------------------
// K5 + K6: per tile, in tet order: final active list (tet, masks, record, output offsets), vertex
// candidates with their canonical keys, and the hash-min insertion of every shareable candidate.
// The reference's try_emplace(key, next id) (src/extract_mesh.cpp:139,169,214) gives a shared vertex the
// id of its FIRST candidate in (tet, local index) order: every candidate atomicMin's its index into the
// open-addressing slot owned by its key, so the slot ends up holding the first candidate whatever the
// execution order.  Keys of other blocks' candidates are read through L2 (ld.cg): they are written,
// fenced and only then published in the table by the same pass.
// ---------------------------------------------------------------------------------------------
struct EmitArgs
{
    const uint4* tets;
    uint32_t R; // != 0: generated grid, the tet's vertex ids are computed (grid_tet)
    FastDiv dR;
    unsigned* tile_ticket;
    const uint32_t* tl_tet;
    const uint32_t* tl_mask;
    const uint32_t* tl_ref;
    size_t tl_stride;
    uint32_t tile_slots;
    const TileTot* tile_tot;
    const TileTot* tile_off;
    uint32_t n_tiles;
    const uint32_t* blob32;
    const uint32_t* arena32;
    uint32_t* act_tet;
    uint32_t* act_mask;
    uint32_t act_cap;
    uint32_t* rec_ref;
    uint4* offs;
    uint4* cand_key;
    uint32_t* cand_src; // candidate -> active tet
    const unsigned* overflow;
};

template <int W>
__global__ void __launch_bounds__(256) emit_kernel(const EmitArgs A)
{
    __shared__ uint4 s_warp[8];
    __shared__ uint4 s_run, s_tot;
    __shared__ unsigned s_tile;
    if (*A.overflow) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        // tiles carry very different numbers of active tets: dynamic assignment
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t;
            do t = atomicAdd(A.tile_ticket, 1u);
            while (t < A.n_tiles && A.tile_tot[t].act == 0);
            s_tile = t;
        }
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= A.n_tiles) return;
        const uint32_t n_act = A.tile_tot[tile].act;
        const TileTot off = A.tile_off[tile];
        if (threadIdx.x == 0) s_run = make_uint4(off.cand, off.face, off.fv, off.funcs);
        __syncthreads();
        for (uint32_t r0 = 0; r0 < n_act; r0 += 256) {
            const uint32_t r = r0 + threadIdx.x;
            const bool valid = r < n_act;
            const size_t slot = (size_t)tile * A.tile_slots + r;
            uint32_t t = 0, ref = 0, m[W];
            const uint32_t* rec = A.blob32;
            uint4 ci = make_uint4(0, 0, 0, 0);
            if (valid) {
                t = A.tl_tet[slot];
                ref = A.tl_ref[slot];
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    m[w] = A.tl_mask[(size_t)w * A.tl_stride + slot];
                    ci.w += __popc(m[w]);
                }
                rec = (ref & REF_GENERAL) ? A.arena32 + (size_t)(ref & ~REF_FLAGS) : A.blob32 + ref;
                const uint32_t h = rec[0];
                ci.x = h & 255;
                ci.y = (h >> 8) & 255;
                ci.z = h >> 16;
            }
            // block-wide exclusive scan of the four counts
            uint4 x = ci;
#pragma unroll
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
                const uint4 tq = (lane < 8) ? s_warp[lane] : make_uint4(0, 0, 0, 0);
                uint4 x8 = tq;
#pragma unroll
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
                if (lane < 8) s_warp[lane] = make_uint4(x8.x - tq.x, x8.y - tq.y, x8.z - tq.z, x8.w - tq.w);
                if (lane == 7) s_tot = x8;
            }
            __syncthreads();
            const uint4 base = s_run, wv = s_warp[warp], tot = s_tot;
            const uint4 o = make_uint4(base.x + wv.x + x.x - ci.x, base.y + wv.y + x.y - ci.y,
                base.z + wv.z + x.z - ci.z, base.w + wv.w + x.w - ci.w);
            if (valid) {
                const uint32_t a = off.act + r;
                A.act_tet[a] = t;
#pragma unroll
                for (int w = 0; w < W; ++w) A.act_mask[(size_t)w * A.act_cap + a] = m[w];
                A.rec_ref[a] = ref;
                A.offs[a] = o;
                if (ci.x) {
                    const uint4 tv4 = A.R ? grid_tet(A.R, A.dR, t) : __ldg(&A.tets[t]);
                    const uint32_t tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
                    uint32_t fl[4];
                    first_funcs<W>(m, fl);
                    for (uint32_t i = 0; i < ci.x; ++i) {
                        A.cand_key[o.x + i] = iso_vert_info<W>(rec[1 + i], tv, fl, m).key;
                        A.cand_src[o.x + i] = a;
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_run = make_uint4(base.x + tot.x, base.y + tot.y, base.z + tot.z, base.w + tot.w);
            __syncthreads();
        }
    }
}

// K6: hash-min insertion, one candidate per thread and INS_BATCH in flight (first probes of all, then the key
// checks).  A full table (sized from the previous pass) is reported through the overflow word.
constexpr int INS_BATCH = 4;
struct InsertArgs
{
    const uint4* cand_key;
    const PassTotals* totals;
    uint32_t* table;
    uint32_t table_mask;
    uint32_t* slot_of;
    unsigned* overflow;
};

__global__ void __launch_bounds__(256) insert_kernel(const InsertArgs A)
{
    if (*A.overflow) return;
    const uint32_t n = A.totals->n_cand;
    const uint32_t step = gridDim.x * 256 * INS_BATCH;
    for (uint32_t c0 = blockIdx.x * 256 * INS_BATCH + threadIdx.x; c0 < n; c0 += step) {
        uint4 k[INS_BATCH];
        uint32_t h[INS_BATCH], cur[INS_BATCH];
        bool act[INS_BATCH];
#pragma unroll
        for (int u = 0; u < INS_BATCH; ++u) {
            const uint32_t c = c0 + u * 256;
            act[u] = c < n;
            if (act[u]) k[u] = A.cand_key[c];
        }
#pragma unroll
        for (int u = 0; u < INS_BATCH; ++u) {
            const uint32_t c = c0 + u * 256;
            if (act[u] && k[u].x == NONE32) { // never shared
                A.slot_of[c] = NONE32;
                act[u] = false;
            }
            if (act[u]) {
                h[u] = hash4(k[u]) & A.table_mask;
                cur[u] = atomicCAS(&A.table[h[u]], NONE32, c);
            }
        }
        uint4 kc[INS_BATCH];
#pragma unroll
        for (int u = 0; u < INS_BATCH; ++u)
            if (act[u] && cur[u] != NONE32) kc[u] = A.cand_key[cur[u]];
#pragma unroll
        for (int u = 0; u < INS_BATCH; ++u) {
            if (!act[u]) continue;
            const uint32_t c = c0 + u * 256;
            uint32_t hh = h[u], cc = cur[u];
            uint4 kk = kc[u];
            for (uint32_t probes = 0;; ++probes) {
                if (cc == NONE32) break; // the slot was empty and is ours now
                if (key_eq(kk, k[u])) {
                    atomicMin(&A.table[hh], c);
                    break;
                }
                if (probes > A.table_mask) {
                    atomicOr(A.overflow, OVF_TABLE);
                    break;
                }
                hh = (hh + 1) & A.table_mask;
                cc = atomicCAS(&A.table[hh], NONE32, c);
                if (cc != NONE32) kk = A.cand_key[cc];
            }
            A.slot_of[c] = hh;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K6 + K7: a candidate is the representative of its key when its slot holds its own index; the id of
// a representative is the number of representatives before it in candidate order (tile look-back scan),
// exactly the reference's first-occurrence numbering.  The representative writes its IsoVert record and
// coordinates (compute_iso_vert_xyz src/extract_mesh.cpp:1446-1538 with compute_barycentric_coords
// src/extract_mesh.h:111-159: same operation order, no FMA contraction) and replaces the table entry
// (or its slot_of entry when it is never shared) by VID_FLAG | id, which is what the face kernel reads.
// ---------------------------------------------------------------------------------------------
struct RankArgs
{
    const uint4* tets;
    uint32_t R;
    FastDiv dR;
    const uint32_t* act_tet;
    const uint32_t* act_mask;
    uint32_t act_cap;
    const uint32_t* rec_ref;
    const uint4* offs;
    const uint32_t* blob32;
    const uint32_t* arena32;
    const PassTotals* totals;
    const uint32_t* cand_src;
    uint32_t* table;
    uint32_t* slot_of;
    const double* vals;
    uint32_t VS;
    const double* pts;
    uint32_t* v_tet;
    uint8_t* v_local;
    uint8_t* v_size;
    uint4* v_simplex;
    uint4* v_funcs;
    double* v_xyz;
    uint4* v_key;
    volatile unsigned long long* status;
    unsigned* tile_counter;
    unsigned* n_unique;
    const unsigned* overflow;
};

__device__ __forceinline__ void iso_vert_xyz(const uint32_t* sv, const uint32_t* fi, int size,
    const double* __restrict__ vals, uint32_t VS, const double* __restrict__ pts, double* out)
{
#define PT(v, c) pts[3 * (size_t)(v) + (c)]
#define FV(v, f) vals[(size_t)(f) * VS + (v)]
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
}

constexpr int RV_ITEMS = 4;
constexpr int RV_TILE = 256 * RV_ITEMS; // candidates per tile

// the representative candidate c of active tet a becomes vertex `id`
template <int W>
__device__ __forceinline__ void write_iso_vertex(const RankArgs& A, uint32_t c, uint32_t s, uint32_t id)
{
    const uint32_t a = A.cand_src[c];
    const uint32_t t = A.act_tet[a];
    const uint32_t ref = A.rec_ref[a];
    const uint32_t i = c - A.offs[a].x;
    uint32_t m[W], fl[4];
#pragma unroll
    for (int w = 0; w < W; ++w) m[w] = A.act_mask[(size_t)w * A.act_cap + a];
    const uint32_t* rec = (ref & REF_GENERAL) ? A.arena32 + (size_t)(ref & ~REF_FLAGS) : A.blob32 + ref;
    const uint32_t e = rec[1 + i];
    const uint4 tv4 = A.R ? grid_tet(A.R, A.dR, t) : __ldg(&A.tets[t]);
    const uint32_t tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
    first_funcs<W>(m, fl);
    const IsoVertInfo vi = iso_vert_info<W>(e, tv, fl, m);
    uint32_t sv[4];
    if (vi.size == 4) {
        sv[0] = tv[0];
        sv[1] = tv[1];
        sv[2] = tv[2];
        sv[3] = tv[3];
    } else {
        sv[0] = vi.key.x;
        sv[1] = vi.key.y;
        sv[2] = vi.key.z;
        sv[3] = NONE32;
    }
    uint32_t fi[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) fi[q] = vi.f[q] == 0xffffu ? NONE32 : vi.f[q];
    A.v_tet[id] = t;
    A.v_local[id] = (uint8_t)vi.local;
    A.v_size[id] = (uint8_t)vi.size;
    A.v_simplex[id] = make_uint4(sv[0], sv[1], sv[2], sv[3]);
    A.v_funcs[id] = make_uint4(fi[0], fi[1], fi[2], NONE32);
    A.v_key[id] = vi.key;
    double out[3];
    iso_vert_xyz(sv, fi, vi.size, A.vals, A.VS, A.pts, out);
    A.v_xyz[3 * (size_t)id + 0] = out[0];
    A.v_xyz[3 * (size_t)id + 1] = out[1];
    A.v_xyz[3 * (size_t)id + 2] = out[2];
    if (s == NONE32)
        A.slot_of[c] = VID_FLAG | id;
    else
        A.table[s] = VID_FLAG | id;
}

template <int W>
__global__ void __launch_bounds__(256, 6) rank_verts_kernel(const RankArgs A)
{
    __shared__ unsigned s_tile, s_base;
    __shared__ unsigned s_cnt[RV_ITEMS * 8];
    __shared__ uint2 s_rep[RV_TILE]; // (candidate, slot) of the tile's representatives, in candidate order
    if (*A.overflow) return;
    const uint32_t n_cand = A.totals->n_cand;
    const uint32_t n_tiles = (n_cand + RV_TILE - 1) / RV_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(A.tile_counter, 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= n_tiles) return;
        // candidates in candidate order: item j of thread x is candidate tile * RV_TILE + j * 256 + x
        uint32_t slot[RV_ITEMS];
        unsigned ball[RV_ITEMS];
        bool rep[RV_ITEMS];
#pragma unroll
        for (int j = 0; j < RV_ITEMS; ++j) {
            const uint32_t c = tile * RV_TILE + j * 256 + threadIdx.x;
            slot[j] = (c < n_cand) ? A.slot_of[c] : 0u;
        }
#pragma unroll
        for (int j = 0; j < RV_ITEMS; ++j) {
            const uint32_t c = tile * RV_TILE + j * 256 + threadIdx.x;
            rep[j] = (c < n_cand) && (slot[j] == NONE32 || __ldcg(&A.table[slot[j]]) == c);
        }
#pragma unroll
        for (int j = 0; j < RV_ITEMS; ++j) {
            ball[j] = __ballot_sync(0xffffffffu, rep[j]);
            if (lane == 0) s_cnt[j * 8 + warp] = __popc(ball[j]);
        }
        __syncthreads();
        // exclusive prefix over the (item, warp) sequence: RV_ITEMS * 8 = 32 entries, one warp scan each
        const unsigned cq = s_cnt[lane];
        unsigned x = cq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        const unsigned run = __shfl_sync(0xffffffffu, x, 31);
        unsigned my_off[RV_ITEMS];
#pragma unroll
        for (int j = 0; j < RV_ITEMS; ++j) my_off[j] = __shfl_sync(0xffffffffu, x - cq, j * 8 + warp);
        if (warp == 0) {
            uint32_t e0, e1;
            tile_lookback_warp(A.status, (int)tile, run, 0, e0, e1);
            if (lane == 0) {
                s_base = e0;
                if (tile == n_tiles - 1) *A.n_unique = e0 + run;
            }
        }
#pragma unroll
        for (int j = 0; j < RV_ITEMS; ++j)
            if (rep[j])
                s_rep[my_off[j] + __popc(ball[j] & ((1u << lane) - 1u))] =
                    make_uint2(tile * RV_TILE + j * 256 + threadIdx.x, slot[j]);
        __syncthreads();
        const unsigned base = s_base;
        // dense: consecutive lanes write consecutive vertices
        for (unsigned i = threadIdx.x; i < run; i += 256) {
            const uint2 q = s_rep[i];
            write_iso_vertex<W>(A, q.x, q.y, base + i);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5 (faces): PolygonFace arrays.  Face ids follow (tet, local face) order = the record order; interior
// iso-faces are never shared (src/extract_mesh.cpp:254-259).  func_index.first =
// func_in_tet[supporting_plane - 4 + start] (:249,:258).
// HDR = true (degenerate inputs only: some iso-face lies on a tet boundary): writes the intermediate
// face headers + vertex lists consumed by the boundary-face matching kernels (bface_*_kernel).
// ---------------------------------------------------------------------------------------------
struct FaceArgs
{
    const uint32_t* act_tet;
    const uint32_t* act_mask;
    uint32_t act_cap;
    const uint32_t* rec_ref;
    const uint4* offs;
    const uint32_t* blob32;
    const uint32_t* arena32;
    const PassTotals* totals;
    const uint32_t* table;
    const uint32_t* slot_of;
    uint32_t* f_off;
    uint32_t* f_verts;
    uint32_t* f_toff;
    uint32_t* f_tets;
    uint32_t* f_funcs;
    uint4* face_hdr; // HDR
    const unsigned* overflow;
};

constexpr int FK_NV = 8; // final ids of a tet's first FK_NV candidates are fetched up front (independent loads)

template <int W, bool HDR>
__global__ void __launch_bounds__(256) faces_kernel(const FaceArgs A)
{
    __shared__ uint32_t s_vid[FK_NV][256];
    if (*A.overflow) return;
    const uint32_t n_active = A.totals->n_active;
    if (!HDR && blockIdx.x == 0 && threadIdx.x == 0) {
        A.f_off[A.totals->n_faces] = A.totals->n_fv;
        A.f_toff[A.totals->n_faces] = A.totals->n_faces;
    }
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_active; a += gridDim.x * blockDim.x) {
        const uint32_t ref = A.rec_ref[a];
        const uint32_t* r = (ref & REF_GENERAL) ? A.arena32 + (size_t)(ref & ~REF_FLAGS) : A.blob32 + ref;
        const uint32_t hdr = r[0];
        const int nv = hdr & 255, nf = (hdr >> 8) & 255;
        if (nf == 0) continue;
        const uint32_t t = A.act_tet[a];
        uint32_t m[W], fl[4];
#pragma unroll
        for (int w = 0; w < W; ++w) m[w] = A.act_mask[(size_t)w * A.act_cap + a];
        first_funcs<W>(m, fl);
        const uint4 o = A.offs[a];
        {
            uint32_t sl[FK_NV];
#pragma unroll
            for (int i = 0; i < FK_NV; ++i) sl[i] = (i < nv) ? A.slot_of[o.x + i] : VID_FLAG;
#pragma unroll
            for (int i = 0; i < FK_NV; ++i)
                s_vid[i][threadIdx.x] = ((sl[i] & VID_FLAG) ? sl[i] : A.table[sl[i]]) & ~VID_FLAG;
        }
        const uint32_t* p = r + 1 + nv;
        uint32_t fvo = o.z;
        for (int j = 0; j < nf; ++j) {
            const uint32_t e = *p++;
            const uint32_t local = e & 0xffffu;
            const int sp = (e >> 16) & 255, n = (e >> 24) & 127, bnd = e >> 31;
            uint32_t f;
            if (sp > 3)
                f = func_of_plane<W>(sp, fl, m);
            else {
                // QUIRK kept from the reference: for a face coplanar with a tet face sp < 4 and the index
                // start + sp - 4 points at a CRS entry of an earlier tet (none: Mesh_None)
                f = NONE32;
                int back = 4 - sp;
                for (uint32_t ap = a; ap > 0 && back > 0;) {
                    --ap;
                    uint32_t mp[W];
                    int kp = 0;
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        mp[w] = A.act_mask[(size_t)w * A.act_cap + ap];
                        kp += __popc(mp[w]);
                    }
                    if (back <= kp) {
                        f = (uint32_t)nth_set_bit(mp, W, kp - back);
                        back = 0;
                    } else
                        back -= kp;
                }
            }
            const uint32_t fi = o.y + j;
            if (HDR)
                A.face_hdr[fi] = make_uint4(t, local | ((uint32_t)n << 16) | ((uint32_t)bnd << 24), f, fvo);
            else {
                A.f_off[fi] = fvo;
                A.f_toff[fi] = fi;
                A.f_tets[2 * (size_t)fi] = t;
                A.f_tets[2 * (size_t)fi + 1] = local;
                A.f_funcs[2 * (size_t)fi] = f;
                A.f_funcs[2 * (size_t)fi + 1] = NONE32;
            }
            for (int k0 = 0; k0 < n; k0 += 4) {
                const uint32_t x = *p++;
                for (int k = k0; k < n && k < k0 + 4; ++k) {
                    const uint32_t rk = (x >> (8 * (k - k0))) & 255;
                    A.f_verts[fvo + k] = rk < FK_NV ? s_vid[rk][threadIdx.x] : final_vid(o.x + rk, A.slot_of, A.table);
                }
            }
            fvo += n;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// K2b on a generated grid: material filter (src/material_interface.cpp:99-152) with the cube structure of the
// implicit-arrangement filter above.  S = union of the corners' maximal materials; a tet is active iff S has at
// least two members (:120-123) - that needs the eight corner masks of the cube only, so the ordered compaction
// is exact after the streaming phase.  The value test (every material exceeding min_h at >= 2 corners joins S,
// :125-145) then runs densely over the tile's active slots.  Output in the layout of filter_mi_tiles_kernel
// (tile-local slots + per-tile (active, CRS length) counts) with MI_GRID_SLOTS slots per tile.
// ---------------------------------------------------------------------------------------------
constexpr int MI_GRID_ROUNDS = 4;
constexpr int MI_GRID_SLOTS = MI_GRID_ROUNDS * 256 * 5;

template <int W>
__global__ void __launch_bounds__(256) filter_mi_grid_kernel(uint32_t R, const FastDiv dR, uint32_t t_first,
    uint32_t t_count, uint32_t c_first, uint32_t n_units, const uint2* __restrict__ vmask,
    const double* __restrict__ vals, uint32_t VS, uint32_t F, uint32_t* __restrict__ tl_tet,
    uint32_t* __restrict__ tl_mask, size_t tl_stride, uint2* __restrict__ tile_cnt, FilterCounters* __restrict__ ctr)
{
    constexpr int ROUNDS = (W == 1) ? MI_GRID_ROUNDS : 1; // more words: fewer cubes per thread in registers
    constexpr int UNITS = MI_GRID_ROUNDS * 256;
    constexpr int NE = MI_GRID_ROUNDS * 8;
    __shared__ unsigned s_cnt[NE];
    __shared__ unsigned s_tot[4];
    const unsigned tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) s_tot[threadIdx.x] = 0;
    const uint32_t t_end = t_first + t_count, N = R + 1;
    const size_t tbase = (size_t)tile * MI_GRID_SLOTS;
    unsigned run_total = 0;
    // W == 1: all four rounds in registers, one barrier.  W > 1: round by round with a running offset.
    for (int r0 = 0; r0 < MI_GRID_ROUNDS; r0 += ROUNDS) {
        uint32_t m[ROUNDS][5][W];
#pragma unroll
        for (int rr = 0; rr < ROUNDS; ++rr) {
            const uint32_t u = tile * UNITS + (r0 + rr) * 256 + threadIdx.x;
            const bool in = u < n_units;
            const uint32_t cube = c_first + (in ? u : 0u);
            const uint32_t ij = fd_div(cube, dR), k = cube - ij * R, i = fd_div(ij, dR), j = ij - i * R;
            const uint32_t base = (i * N + j) * N + k;
            const bool odd = (i + j + k) & 1;
            uint32_t cv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                cv[q] = base + (((q & 3) == 1 || (q & 3) == 2) ? N * N : 0u) + (((q & 3) >= 2) ? N : 0u) + (q >> 2);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t c8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) c8[q] = __ldg(&vmask[(size_t)w * VS + cv[q]]).x;
#define RIN_TS(a, b, c, d) (c8[a] | c8[b] | c8[c] | c8[d])
                uint32_t x[5];
                if (!odd) {
                    x[0] = RIN_TS(4, 6, 1, 3);
                    x[1] = RIN_TS(6, 3, 4, 7);
                    x[2] = RIN_TS(1, 3, 0, 4);
                    x[3] = RIN_TS(3, 1, 2, 6);
                    x[4] = RIN_TS(4, 1, 6, 5);
                } else {
                    x[0] = RIN_TS(7, 0, 2, 5);
                    x[1] = RIN_TS(2, 3, 0, 7);
                    x[2] = RIN_TS(5, 7, 0, 4);
                    x[3] = RIN_TS(7, 2, 6, 5);
                    x[4] = RIN_TS(0, 1, 2, 5);
                }
#undef RIN_TS
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    const uint32_t t = 5 * cube + s;
                    m[rr][s][w] = (!in || t < t_first || t >= t_end) ? 0u : x[s];
                }
            }
        }
        // active iff the union has >= 2 members
        unsigned pr[ROUNDS];
        uint32_t act[ROUNDS];
#pragma unroll
        for (int rr = 0; rr < ROUNDS; ++rr) {
            unsigned cnt = 0;
            act[rr] = 0;
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                int ns = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) ns += __popc(m[rr][s][w]);
                if (ns >= 2) {
                    act[rr] |= 1u << s;
                    ++cnt;
                }
            }
            unsigned x = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            pr[rr] = x - cnt;
            if (lane == 31) s_cnt[(r0 + rr) * 8 + warp] = x;
        }
        __syncthreads();
        // prefix over the (round, warp) entries written so far in this pass: rounds r0 .. r0 + ROUNDS - 1
        unsigned my_off[ROUNDS];
        {
            const int e = lane; // NE == 32
            const bool mine = e >= r0 * 8 && e < (r0 + ROUNDS) * 8;
            const unsigned c = mine ? s_cnt[e] : 0u;
            unsigned x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
#pragma unroll
            for (int rr = 0; rr < ROUNDS; ++rr)
                my_off[rr] = run_total + __shfl_sync(0xffffffffu, x - c, (r0 + rr) * 8 + warp);
            run_total += __shfl_sync(0xffffffffu, x, 31);
        }
#pragma unroll
        for (int rr = 0; rr < ROUNDS; ++rr) {
            if (s_cnt[(r0 + rr) * 8 + warp] == 0) continue;
            unsigned rank = 0;
#pragma unroll
            for (int s = 0; s < 5; ++s)
                if ((act[rr] >> s) & 1) {
                    const uint32_t u = tile * UNITS + (r0 + rr) * 256 + threadIdx.x;
                    const size_t pos = tbase + my_off[rr] + pr[rr] + rank++;
                    tl_tet[pos] = 5 * (c_first + u) + s;
#pragma unroll
                    for (int w = 0; w < W; ++w) tl_mask[(size_t)w * tl_stride + pos] = m[rr][s][w];
                }
        }
        __syncthreads();
    }
    // ---- value test, dense over the tile's active slots
    unsigned k1 = 0, k2 = 0, km = 0, kf = 0;
    for (uint32_t i = threadIdx.x; i < run_total; i += 256) {
        const size_t pos = tbase + i;
        const uint4 tv = grid_tet(R, dR, __ldcg(&tl_tet[pos]));
        const uint32_t vv[4] = {tv.x, tv.y, tv.z, tv.w};
        uint32_t mw[W];
#pragma unroll
        for (int w = 0; w < W; ++w) mw[w] = __ldcg(&tl_mask[(size_t)w * tl_stride + pos]);
        double min_h[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) min_h[c] = 1.7976931348623157e308;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            uint32_t sset = mw[w];
            while (sset) {
                const int f = w * 32 + __ffs(sset) - 1;
                sset &= sset - 1;
#pragma unroll
                for (int c = 0; c < 4; ++c) min_h[c] = fmin(min_h[c], __ldg(&vals[(size_t)f * VS + vv[c]]));
            }
        }
        int kq = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            for (uint32_t f = w * 32; f < F && f < (uint32_t)w * 32 + 32; ++f) {
                int greater = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) greater += (__ldg(&vals[(size_t)f * VS + vv[c]]) > min_h[c]);
                if (greater > 1) mw[w] |= 1u << (f & 31);
            }
            tl_mask[(size_t)w * tl_stride + pos] = mw[w];
            kq += __popc(mw[w]);
        }
        k1 += (kq == 2);
        k2 += (kq == 3);
        km += (kq > 3);
        kf += kq;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        k1 += __shfl_xor_sync(0xffffffffu, k1, o);
        k2 += __shfl_xor_sync(0xffffffffu, k2, o);
        km += __shfl_xor_sync(0xffffffffu, km, o);
        kf += __shfl_xor_sync(0xffffffffu, kf, o);
    }
    if (lane == 0) {
        if (k1) atomicAdd(&s_tot[0], k1);
        if (k2) atomicAdd(&s_tot[1], k2);
        if (km) atomicAdd(&s_tot[2], km);
        if (kf) atomicAdd(&s_tot[3], kf);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tile_cnt[tile] = make_uint2(run_total, s_tot[3]);
        if (s_tot[0]) atomicAdd(&ctr->n_k1, s_tot[0]);
        if (s_tot[1]) atomicAdd(&ctr->n_k2, s_tot[1]);
        if (s_tot[2]) atomicAdd(&ctr->n_kmore, s_tot[2]);
    }
}

} // namespace rin
