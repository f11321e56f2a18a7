// Device-side slab-boundary exchange over NCCL (NVLink / NVSwitch): the whole protocol of
// sharding.py as stream-ordered kernels around two ncclAllGather calls, no host round trips
// except the final read-back of the counts.
#pragma once
#include "kernels_ia.cuh"

namespace rin {

// message layout per rank (uint32 words): header[XHDR] = {count, n_own, n_faces, n_face_verts, n_face_tets, 0..},
// keys[cap][4], ids[cap]
constexpr uint32_t XHDR = 8;
__host__ __device__ inline size_t xmsg_words(uint32_t cap)
{
    return XHDR + (size_t)cap * 5;
}

__global__ void x_header_kernel(uint32_t* msg, const unsigned* count, uint32_t cap, const unsigned* n_own,
    uint32_t n_faces, unsigned* overflow, uint32_t n_fv = 0, uint32_t n_ft = 0, uint32_t need_ghost = 0)
{
    const unsigned c = *count;
    msg[0] = c;
    msg[1] = n_own ? *n_own : 0;
    msg[2] = n_faces;
    msg[3] = n_fv;
    msg[4] = n_ft;
    msg[5] = need_ghost; // degenerate vertices seen by a run without ghost tets
    msg[6] = msg[7] = 0;
    if (c > cap) *overflow = c;
}

// select into a message buffer (keys at msg+4, ids at msg+4+4*cap)
__global__ void __launch_bounds__(256) x_select_kernel(const uint4* __restrict__ v_key,
    const uint8_t* __restrict__ v_size, uint32_t n, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ own_idx,
    uint32_t* __restrict__ msg, uint32_t cap, unsigned* __restrict__ n_out, const unsigned* __restrict__ n_dev = nullptr)
{
    if (n_dev) n = *n_dev; // fused run + exchange: the vertex count has not reached the host yet
    uint4* out_keys = reinterpret_cast<uint4*>(msg + XHDR);
    uint32_t* out_ids = msg + XHDR + (size_t)cap * 4;
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

// inserts the keys of ranks [0, rank) of a gathered buffer into the foreign table;
// table[h] = (s * cap + i) indexes the gathered buffer
__global__ void __launch_bounds__(256) x_insert_kernel(const uint32_t* __restrict__ all, uint32_t cap, int rank,
    uint32_t* __restrict__ table, uint32_t mask)
{
    const size_t stride = xmsg_words(cap);
    const uint32_t total = (uint32_t)rank * cap;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
        const uint32_t s = c / cap, i = c % cap;
        const uint32_t* m = all + s * stride;
        if (i >= min(m[0], cap)) continue;
        const uint4 k = reinterpret_cast<const uint4*>(m + XHDR)[i];
        uint32_t h = hash4(k) & mask;
        for (;;) {
            const uint32_t cur = atomicCAS(&table[h], NONE32, c);
            if (cur == NONE32) break;
            const uint4 kc = reinterpret_cast<const uint4*>(all + (cur / cap) * stride + XHDR)[cur % cap];
            if (key_eq(kc, k)) break;
            h = (h + 1) & mask;
        }
    }
}

__device__ __forceinline__ uint32_t x_lookup(const uint32_t* all, uint32_t cap, const uint32_t* table, uint32_t mask,
    uint4 k)
{
    const size_t stride = xmsg_words(cap);
    uint32_t h = hash4(k) & mask;
    for (;;) {
        const uint32_t cur = table[h];
        if (cur == NONE32) return NONE32;
        const uint4 kc = reinterpret_cast<const uint4*>(all + (cur / cap) * stride + XHDR)[cur % cap];
        if (key_eq(kc, k)) return cur;
        h = (h + 1) & mask;
    }
}

// own flag + ordered own index (decoupled look-back over tiles of 1024 vertices)
__global__ void __launch_bounds__(256) x_mark_scan_kernel(const uint4* __restrict__ v_key,
    const uint8_t* __restrict__ v_size, uint32_t n, const uint32_t* __restrict__ all, uint32_t cap,
    const uint32_t* __restrict__ table, uint32_t mask, int rank, uint32_t* __restrict__ own_idx,
    volatile unsigned long long* __restrict__ status, unsigned* __restrict__ tile_counter,
    unsigned* __restrict__ n_own, uint32_t v_lo, const unsigned* __restrict__ n_dev = nullptr)
{
    __shared__ unsigned s_tile, s_base;
    __shared__ unsigned s_warp[8];
    if (n_dev) n = *n_dev;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if ((size_t)tile * 1024 >= n) return; // launched for the capacity: no later tile depends on this one
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ITEMS = 4;
    const uint32_t base = tile * 256 * ITEMS + threadIdx.x * ITEMS;
    bool own[ITEMS];
    unsigned cnt = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t i = base + j;
        own[j] = false;
        if (i < n) {
            own[j] = true;
            if (rank > 0 && v_size[i] < 4 && x_lookup(all, cap, table, mask, v_key[i]) != NONE32) own[j] = false;
            if (i < v_lo) own[j] = false; // first created by a ghost tet of the rank below
            cnt += own[j];
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
            if (tile == (n + 256 * ITEMS - 1) / (256 * ITEMS) - 1) *n_own = e0 + run;
        }
    }
    __syncthreads();
    unsigned id = s_base + s_warp[warp] + x - cnt;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const uint32_t i = base + j;
        if (i < n) own_idx[i] = own[j] ? id++ : NONE32;
    }
}

// offsets[s] = sum of n_own of ranks < s; offsets[world] = total; face offsets likewise
// `first` is the gathered buffer of the first collective: every rank derives the SAME overflow
// decision from the same gathered headers (a rank deciding alone would desynchronise the collectives)
__global__ void x_offsets_kernel(const uint32_t* __restrict__ first, const uint32_t* __restrict__ all, uint32_t cap,
    int world, uint32_t* __restrict__ voff, uint32_t* __restrict__ foff, unsigned* __restrict__ overflow,
    unsigned* __restrict__ need_ghost)
{
    if (threadIdx.x || blockIdx.x) return;
    const size_t stride = xmsg_words(cap);
    uint32_t v = 0, f = 0, fv = 0, ft = 0;
    uint32_t* fvoff = foff + (world + 1);
    uint32_t* ftoff = fvoff + (world + 1);
    unsigned ovf = 0, ng = 0;
    for (int s = 0; s < world; ++s) {
        voff[s] = v;
        foff[s] = f;
        fvoff[s] = fv;
        ftoff[s] = ft;
        const uint32_t* m = all + s * stride;
        v += m[1];
        f += m[2];
        fv += m[3];
        ft += m[4];
        ng |= m[5];
        if (m[0] > cap) ovf = max(ovf, m[0]);
        if (first[s * stride] > cap) ovf = max(ovf, first[s * stride]);
    }
    *overflow = ovf;
    *need_ghost = ng;
    voff[world] = v;
    foff[world] = f;
    fvoff[world] = fv;
    ftoff[world] = ft;
}

// global id of every local vertex: own -> offset + own index; foreign -> owner's offset + its own index
__global__ void __launch_bounds__(256) x_global_ids_kernel(const uint4* __restrict__ v_key,
    const uint32_t* __restrict__ own_idx, uint32_t n, int rank, const uint32_t* __restrict__ voff,
    const uint32_t* __restrict__ all, uint32_t cap, const uint32_t* __restrict__ table, uint32_t mask,
    uint32_t* __restrict__ gid, unsigned* __restrict__ n_unresolved, uint32_t v_lo)
{
    const size_t stride = xmsg_words(cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t o = own_idx[i];
        if (o != NONE32) {
            gid[i] = voff[rank] + o;
            continue;
        }
        const uint32_t f = x_lookup(all, cap, table, mask, v_key[i]);
        if (f == NONE32) {
            if (i >= v_lo) atomicAdd(n_unresolved, 1u); // a ghost's vertex off the shared plane: no kept face uses it
            gid[i] = NONE32;
        } else {
            const uint32_t s = f / cap, j = f % cap;
            gid[i] = voff[s] + (all + s * stride + XHDR + (size_t)cap * 4)[j];
        }
    }
}

__global__ void __launch_bounds__(256) add_offset_kernel(uint32_t* __restrict__ a, uint32_t n, uint32_t off)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] += off;
}

// ---- neighbour protocol (slab sharding: a rank shares vertices with ranks r-1 and r+1 only) ----------------
// message layout as above: header[4] = {count, n_own, n_faces, 0}, keys[cap][4], ids[cap]

// own indices of the vertices this rank sent upwards (it owns all of them: it is the lowest rank on that plane)
// out = this rank's record of the count exchange: 8 header words {n_own, n_faces, candidates sent upwards,
// n_face_verts, n_face_tets, 0, 0, 0} followed by the own indices (one all-gather moves both)
__global__ void __launch_bounds__(256) x_own_ids_kernel(const uint32_t* __restrict__ sent, uint32_t cap,
    const uint32_t* __restrict__ own_idx, uint32_t* __restrict__ out, unsigned* __restrict__ n_bad,
    const unsigned* __restrict__ n_own, uint32_t n_faces, uint32_t n_fv, uint32_t n_ft, uint32_t need_ghost)
{
    const uint32_t n = min(sent[0], cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = *n_own;
        out[1] = n_faces;
        out[2] = sent[0];
        out[3] = n_fv;
        out[4] = n_ft;
        out[5] = need_ghost; // degenerate vertices seen by a run without ghost tets
        out[6] = out[7] = 0;
    }
    const uint32_t* ids = sent + XHDR + (size_t)cap * 4;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t o = own_idx[ids[j]];
        if (o == NONE32) atomicAdd(n_bad, 1u);
        out[8 + j] = o;
    }
}

// prefixes of the gathered records (vertices, faces, face-vertex entries, face-tet pairs: four arrays of
// world + 1 entries behind each other); overflow = largest n_up beyond the capacity
__global__ void x_offsets_nb_kernel(const uint32_t* __restrict__ all, size_t stride, int world, uint32_t cap,
    uint32_t* __restrict__ voff, uint32_t* __restrict__ foff, unsigned* __restrict__ overflow,
    unsigned* __restrict__ need_ghost)
{
    if (threadIdx.x || blockIdx.x) return;
    uint32_t v = 0, f = 0, fv = 0, ft = 0;
    uint32_t* fvoff = foff + (world + 1);
    uint32_t* ftoff = fvoff + (world + 1);
    unsigned ovf = 0, ng = 0, bad = 0;
    for (int s = 0; s < world; ++s) {
        voff[s] = v;
        foff[s] = f;
        fvoff[s] = fv;
        ftoff[s] = ft;
        v += all[stride * s];
        f += all[stride * s + 1];
        fv += all[stride * s + 3];
        ft += all[stride * s + 4];
        ng |= all[stride * s + 5];
        bad |= all[stride * s + 6];
        if (all[stride * s + 2] > cap) ovf = max(ovf, all[stride * s + 2]);
    }
    *need_ghost = ng;
    need_ghost[1] = bad; // fused run + exchange: some rank's run has to be repeated
    voff[world] = v;
    foff[world] = f;
    fvoff[world] = fv;
    ftoff[world] = ft;
    *overflow = ovf;
}

// global ids: own -> offset + own index; foreign -> offset of the lower neighbour + its own index
__global__ void __launch_bounds__(256) x_global_ids_nb_kernel(const uint4* __restrict__ v_key,
    const uint32_t* __restrict__ own_idx, uint32_t n, int rank, const uint32_t* __restrict__ voff,
    const uint32_t* __restrict__ recv_low, const uint32_t* __restrict__ ids_low, uint32_t cap,
    const uint32_t* __restrict__ table, uint32_t mask, uint32_t* __restrict__ gid, unsigned* __restrict__ n_unresolved,
    uint32_t v_lo, const unsigned* __restrict__ n_dev = nullptr)
{
    if (n_dev) n = *n_dev;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t o = own_idx[i];
        if (o != NONE32) {
            gid[i] = voff[rank] + o;
            continue;
        }
        const uint32_t f = (rank > 0) ? x_lookup(recv_low, cap, table, mask, v_key[i]) : NONE32;
        if (f == NONE32 || ids_low[f] == NONE32) {
            if (i >= v_lo) atomicAdd(n_unresolved, 1u);
            gid[i] = NONE32;
        } else
            gid[i] = voff[rank - 1] + ids_low[f];
    }
}

// ---- ghost tets (sharded runs on degenerate inputs, rin_set_ghost_tets) ------------------------------------
// Vertices are numbered in the order of their creating tet and faces are listed in the order of theirs (first
// entry of the tet list), so what the ghosts below / above created is a prefix / suffix of both arrays.
struct GhostBounds
{
    uint32_t v_lo, v_hi;   // vertices created by own tets: [v_lo, v_hi)
    uint32_t f_lo, f_hi;   // faces created by own tets
    uint32_t fv_lo, fv_hi; // their entries in the face-vertex array
    uint32_t ft_lo, ft_hi; // and in the (tet, local face) array
};

__global__ void ghost_bounds_kernel(const uint32_t* __restrict__ v_tet, uint32_t nv, const uint32_t* __restrict__ f_off,
    const uint32_t* __restrict__ f_toff, const uint32_t* __restrict__ f_tets, uint32_t nf, uint32_t own_first,
    uint32_t own_end, GhostBounds* __restrict__ out)
{
    if (threadIdx.x || blockIdx.x) return;
    auto first_vert = [&](uint32_t t) { // first vertex whose creating tet is >= t
        uint32_t lo = 0, hi = nv;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (v_tet[mid] < t) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    auto first_face = [&](uint32_t t) {
        uint32_t lo = 0, hi = nf;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (f_tets[2 * (size_t)f_toff[mid]] < t) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    GhostBounds b;
    b.v_lo = first_vert(own_first);
    b.v_hi = first_vert(own_end);
    b.f_lo = first_face(own_first);
    b.f_hi = first_face(own_end);
    b.fv_lo = f_off[b.f_lo];
    b.fv_hi = f_off[b.f_hi];
    b.ft_lo = f_toff[b.f_lo];
    b.ft_hi = f_toff[b.f_hi];
    *out = b;
}

// faces [f_lo, f_hi) with their offsets rebased to 0
__global__ void __launch_bounds__(256) ghost_slice_faces_kernel(const GhostBounds b, const uint32_t* __restrict__ f_off,
    const uint32_t* __restrict__ f_verts, const uint32_t* __restrict__ f_toff, const uint32_t* __restrict__ f_tets,
    const uint32_t* __restrict__ f_funcs, uint32_t* __restrict__ o_off, uint32_t* __restrict__ o_verts,
    uint32_t* __restrict__ o_toff, uint32_t* __restrict__ o_tets, uint32_t* __restrict__ o_funcs)
{
    const uint32_t nf = b.f_hi - b.f_lo, nfv = b.fv_hi - b.fv_lo, nft = b.ft_hi - b.ft_lo;
    const uint32_t step = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = i0; i <= nf; i += step) {
        o_off[i] = f_off[b.f_lo + i] - b.fv_lo;
        o_toff[i] = f_toff[b.f_lo + i] - b.ft_lo;
    }
    for (uint32_t i = i0; i < 2 * nf; i += step) o_funcs[i] = f_funcs[2 * (size_t)b.f_lo + i];
    for (uint32_t i = i0; i < nfv; i += step) o_verts[i] = f_verts[b.fv_lo + i];
    for (uint32_t i = i0; i < 2 * nft; i += step) o_tets[i] = f_tets[2 * (size_t)b.ft_lo + i];
}

// ---- peer-memory exchange (rin_run_exchange on one NVLink / NVSwitch node) --------------------------------
// The same neighbour protocol without NCCL on the data path: every rank owns an INBOX in its HBM that the other
// ranks map through CUDA IPC; a sender stores its message straight into the receiver's inbox over NVLink, fences
// (system scope) and then stores the pass number into the receiver's flag word; the receiving kernel spins on that
// flag.  Lower rank -> upper rank for the keys (no cycle), everyone -> everyone for the 8-word count records (a rank
// publishes before it waits).  Inboxes are double-buffered by the parity of the pass number: a rank can be at most
// one pass ahead of a peer, because finishing a pass needs every peer's record of that pass.
//
// inbox layout (uint32 words): flags[PX_FLAG_WORDS] | per parity: keys message xmsg_words(cap) | world records of
// rec_words = 8 + cap
constexpr uint32_t PX_FLAG_WORDS = 128; // [parity * 64 + 0] keys flag, [parity * 64 + 1 + s] record flag of rank s
constexpr int PX_MAX_WORLD = 32;
struct PeerInboxes
{
    uint32_t* p[PX_MAX_WORLD];
};
__host__ __device__ inline size_t px_parity_words(uint32_t cap, int world)
{
    return xmsg_words(cap) + (size_t)world * (8 + (size_t)cap);
}
__host__ __device__ inline size_t px_inbox_words(uint32_t cap, int world)
{
    return PX_FLAG_WORDS + 2 * px_parity_words(cap, world);
}
__host__ __device__ __forceinline__ uint32_t* px_keys(uint32_t* inbox, uint32_t cap, int world, uint32_t parity)
{
    return inbox + PX_FLAG_WORDS + parity * px_parity_words(cap, world);
}
__host__ __device__ __forceinline__ uint32_t* px_record(uint32_t* inbox, uint32_t cap, int world, uint32_t parity, int s)
{
    return px_keys(inbox, cap, world, parity) + xmsg_words(cap) + (size_t)s * (8 + (size_t)cap);
}
__device__ __forceinline__ void px_signal(uint32_t* flag, uint32_t pass)
{
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(flag) = pass;
}
// spins until the flag reaches the pass number; gives up after ~2 s (a peer died) and reports through *timeout
__device__ __forceinline__ void px_wait(const uint32_t* flag, uint32_t pass, unsigned* timeout)
{
    const long long t0 = clock64();
    while ((int32_t)(*reinterpret_cast<const volatile uint32_t*>(flag) - pass) < 0) {
        if (clock64() - t0 > 4000000000ll) {
            atomicExch(timeout, 1u);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

// keys of the vertices on the plane shared with rank + 1, stored straight into its inbox; their local ids stay here
__global__ void __launch_bounds__(256) px_send_keys_kernel(const uint4* __restrict__ v_key,
    const uint8_t* __restrict__ v_size, const unsigned* __restrict__ n_dev, uint32_t lo, uint32_t hi,
    uint32_t* __restrict__ remote_inbox, uint32_t cap, int world, uint32_t pass, uint32_t* __restrict__ local_ids,
    unsigned* __restrict__ n_out, unsigned* __restrict__ done, unsigned* __restrict__ overflow)
{
    const uint32_t n = *n_dev;
    uint32_t* msg = remote_inbox ? px_keys(remote_inbox, cap, world, pass & 1u) : nullptr;
    uint4* out_keys = reinterpret_cast<uint4*>(msg + XHDR);
    bool wrote = false;
    if (msg && lo <= hi)
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int sz = v_size[i];
            if (sz >= 4) continue;
            const uint4 k = v_key[i];
            bool in = k.x >= lo && k.x <= hi;
            if (sz >= 2) in &= k.y >= lo && k.y <= hi;
            if (sz >= 3) in &= k.z >= lo && k.z <= hi;
            if (!in) continue;
            const unsigned p = atomicAdd(n_out, 1u);
            if (p < cap) {
                out_keys[p] = k;
                local_ids[p] = i;
                wrote = true;
            }
        }
    // last block: header, then the flag (the stores of every thread that wrote are fenced before its block's ticket;
    // a system-scope fence is expensive, most threads have nothing to fence)
    if (wrote) __threadfence_system();
    __syncthreads();
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last || threadIdx.x) return;
    const unsigned c = *reinterpret_cast<volatile unsigned*>(n_out);
    if (c > cap) *overflow = c;
    if (msg) {
        msg[0] = c;
        px_signal(remote_inbox + (pass & 1u) * 64, pass);
    }
}

// waits for the lower neighbour's keys of this pass, then inserts them into the foreign table
__global__ void __launch_bounds__(256) px_recv_insert_kernel(uint32_t* __restrict__ inbox, uint32_t cap, int world,
    uint32_t pass, uint32_t* __restrict__ table, uint32_t mask, unsigned* __restrict__ timeout)
{
    if (threadIdx.x == 0) px_wait(inbox + (pass & 1u) * 64, pass, timeout);
    __syncthreads();
    const uint32_t* m = px_keys(inbox, cap, world, pass & 1u);
    const uint32_t n = min(__ldcg(m), cap);
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        const uint4 k = __ldcg(reinterpret_cast<const uint4*>(m + XHDR) + c);
        uint32_t h = hash4(k) & mask;
        for (;;) {
            const uint32_t cur = atomicCAS(&table[h], NONE32, c);
            if (cur == NONE32) break;
            const uint4 kc = __ldcg(reinterpret_cast<const uint4*>(m + XHDR) + cur);
            if (key_eq(kc, k)) break;
            h = (h + 1) & mask;
        }
    }
}

// this rank's count record {n_own, n_faces, n_up, n_fv, n_ft, degenerate, run incomplete, 0} into every inbox, the
// own indices of the vertices sent upwards into the upper neighbour's (and this rank's own) copy, then the flags
__global__ void __launch_bounds__(256) px_publish_record_kernel(const PeerInboxes peers, int rank, int world,
    uint32_t cap, uint32_t pass, const uint32_t* __restrict__ local_ids, const unsigned* __restrict__ n_up,
    const uint32_t* __restrict__ own_idx, const unsigned* __restrict__ n_own, const unsigned* __restrict__ n_faces,
    const unsigned* __restrict__ n_fv, const unsigned long long* __restrict__ n_zero, uint32_t known_degenerate,
    const unsigned* __restrict__ run_overflow, const unsigned* __restrict__ gen_err,
    const unsigned* __restrict__ gen_arena_overflow, const unsigned* __restrict__ n_bnd_faces,
    unsigned* __restrict__ n_bad, unsigned* __restrict__ done)
{
    const uint32_t parity = pass & 1u;
    const uint32_t sent = *n_up, n = min(sent, cap);
    uint32_t* up_rec = (rank + 1 < world) ? px_record(peers.p[rank + 1], cap, world, parity, rank) : nullptr;
    if (up_rec)
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
            const uint32_t o = own_idx[local_ids[j]];
            if (o == NONE32) atomicAdd(n_bad, 1u);
            up_rec[8 + j] = o;
        }
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if ((int)threadIdx.x < world) {
        uint32_t* rec = px_record(peers.p[threadIdx.x], cap, world, parity, rank);
        rec[0] = *n_own;
        rec[1] = *n_faces;
        rec[2] = sent;
        rec[3] = *n_fv;
        rec[4] = *n_faces; // interior iso-faces: one (tet, local face) pair each
        rec[5] = (known_degenerate || *n_zero != 0 || *n_bnd_faces != 0) ? 1u : 0u;
        rec[6] = (*run_overflow || *gen_err || *gen_arena_overflow) ? 1u : 0u;
        rec[7] = 0;
        px_signal(peers.p[threadIdx.x] + parity * 64 + 1 + rank, pass);
    }
}

// px_publish_record_kernel's work as a device function for ONE block (the last block of the mark + scan kernel calls
// it: the sent vertices are a few thousand at most)
struct PublishArgs
{
    PeerInboxes peers;
    int rank, world;
    uint32_t cap, pass;
    const uint32_t* local_ids;
    const unsigned* n_up;
    const unsigned* n_faces;
    const unsigned* n_fv;
    const unsigned long long* n_zero;
    uint32_t known_degenerate;
    const unsigned* run_overflow;
    const unsigned* gen_err;
    const unsigned* gen_arena_overflow;
    const unsigned* n_bnd_faces;
    unsigned* n_bad;
};
__device__ void px_publish_block(const PublishArgs& A, const uint32_t* __restrict__ own_idx, unsigned n_own)
{
    const uint32_t parity = A.pass & 1u;
    const uint32_t sent = *reinterpret_cast<const volatile unsigned*>(A.n_up), n = min(sent, A.cap);
    uint32_t* up_rec = (A.rank + 1 < A.world) ? px_record(A.peers.p[A.rank + 1], A.cap, A.world, parity, A.rank) : nullptr;
    bool wrote = false;
    if (up_rec)
        for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
            const uint32_t o = __ldcg(&own_idx[A.local_ids[j]]);
            if (o == NONE32) atomicAdd(A.n_bad, 1u);
            up_rec[8 + j] = o;
            wrote = true;
        }
    if (wrote) __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < A.world) {
        uint32_t* rec = px_record(A.peers.p[threadIdx.x], A.cap, A.world, parity, A.rank);
        rec[0] = n_own;
        rec[1] = *A.n_faces;
        rec[2] = sent;
        rec[3] = *A.n_fv;
        rec[4] = *A.n_faces; // interior iso-faces: one (tet, local face) pair each
        rec[5] = (A.known_degenerate || *A.n_zero != 0 || *A.n_bnd_faces != 0) ? 1u : 0u;
        rec[6] = (*A.run_overflow || *A.gen_err || *A.gen_arena_overflow) ? 1u : 0u;
        rec[7] = 0;
        px_signal(A.peers.p[threadIdx.x] + parity * 64 + 1 + A.rank, A.pass);
    }
}

// own flag + ordered own index like x_mark_scan_kernel, persistent blocks taking tile tickets (the look-back only
// ever waits for lower tickets, which resident blocks hold); the block that finishes last publishes the record
__global__ void __launch_bounds__(256) px_mark_scan_publish_kernel(const uint4* __restrict__ v_key,
    const uint8_t* __restrict__ v_size, const unsigned* __restrict__ n_dev, const uint32_t* __restrict__ low,
    uint32_t cap, const uint32_t* __restrict__ table, uint32_t mask, int has_low, uint32_t* __restrict__ own_idx,
    volatile unsigned long long* __restrict__ status, unsigned* __restrict__ tile_counter,
    unsigned* __restrict__ n_own, unsigned* __restrict__ done, const PublishArgs P)
{
    __shared__ unsigned s_tile, s_base, s_last;
    __shared__ unsigned s_warp[8];
    const uint32_t n = *n_dev;
    const uint32_t n_tiles = (n + 1023) / 1024;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ITEMS = 4;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const unsigned tile = s_tile;
        if (tile >= n_tiles) break;
        const uint32_t base = tile * 256 * ITEMS + threadIdx.x * ITEMS;
        bool own[ITEMS];
        unsigned cnt = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t i = base + j;
            own[j] = false;
            if (i < n) {
                own[j] = true;
                if (has_low && v_size[i] < 4 && x_lookup(low, cap, table, mask, v_key[i]) != NONE32) own[j] = false;
                cnt += own[j];
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
                if (tile == n_tiles - 1) *n_own = e0 + run;
            }
        }
        __syncthreads();
        unsigned id = s_base + s_warp[warp] + x - cnt;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t i = base + j;
            if (i < n) own_idx[i] = own[j] ? id++ : NONE32;
        }
    }
    // every block signs off once its tiles are written; the last one publishes
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    px_publish_block(P, own_idx, *reinterpret_cast<volatile unsigned*>(n_own));
}

// waits for every rank's record of this pass, computes the prefixes (every block for itself, block 0 also stores
// them for the host), then global ids and the compaction of the owned vertices in one sweep
__global__ void __launch_bounds__(256) px_finish_kernel(uint32_t* __restrict__ inbox, int rank, int world, uint32_t cap,
    uint32_t pass, uint32_t* __restrict__ voff, uint32_t* __restrict__ foff, unsigned* __restrict__ flags /* small */,
    const unsigned* __restrict__ n_dev, const uint4* __restrict__ v_key, const uint32_t* __restrict__ own_idx,
    const uint32_t* __restrict__ table, uint32_t mask, uint32_t* __restrict__ gid,
    const uint32_t* __restrict__ v_tet, const uint8_t* __restrict__ v_local, const uint8_t* __restrict__ v_size,
    const uint4* __restrict__ v_simplex, const uint4* __restrict__ v_funcs, const double* __restrict__ v_xyz,
    uint32_t* __restrict__ o_tet, uint8_t* __restrict__ o_local, uint8_t* __restrict__ o_size,
    uint4* __restrict__ o_simplex, uint4* __restrict__ o_funcs, double* __restrict__ o_xyz, uint4* __restrict__ o_key)
{
    const uint32_t parity = pass & 1u;
    __shared__ uint32_t s_voff[PX_MAX_WORLD + 1];
    if ((int)threadIdx.x < world) px_wait(inbox + parity * 64 + 1 + threadIdx.x, pass, flags + 10);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t v = 0, f = 0, fv = 0, ft = 0;
        unsigned ovf = 0, ng = 0, bad = 0;
        const bool store = blockIdx.x == 0;
        uint32_t* fvoff = foff + (world + 1);
        uint32_t* ftoff = fvoff + (world + 1);
        for (int s = 0; s < world; ++s) {
            const uint32_t* r = px_record(inbox, cap, world, parity, s);
            s_voff[s] = v;
            if (store) {
                voff[s] = v;
                foff[s] = f;
                fvoff[s] = fv;
                ftoff[s] = ft;
            }
            v += __ldcg(r + 0);
            f += __ldcg(r + 1);
            fv += __ldcg(r + 3);
            ft += __ldcg(r + 4);
            ng |= __ldcg(r + 5);
            bad |= __ldcg(r + 6);
            if (__ldcg(r + 2) > cap) ovf = max(ovf, __ldcg(r + 2));
        }
        s_voff[world] = v;
        if (store) {
            voff[world] = v;
            foff[world] = f;
            fvoff[world] = fv;
            ftoff[world] = ft;
            if (ovf > flags[3]) flags[3] = ovf;
            flags[6] = ng;
            flags[7] = bad;
        }
    }
    __syncthreads();
    const uint32_t n = *n_dev;
    const uint32_t* low = px_keys(inbox, cap, world, parity);
    const uint32_t* ids_low = px_record(inbox, cap, world, parity, max(rank - 1, 0)) + 8;
    const uint32_t my_off = s_voff[rank], low_off = s_voff[max(rank - 1, 0)];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t d = own_idx[i];
        if (d != NONE32) {
            gid[i] = my_off + d;
            o_tet[d] = v_tet[i];
            o_local[d] = v_local[i];
            o_size[d] = v_size[i];
            o_simplex[d] = v_simplex[i];
            o_funcs[d] = v_funcs[i];
            o_xyz[3 * (size_t)d] = v_xyz[3 * (size_t)i];
            o_xyz[3 * (size_t)d + 1] = v_xyz[3 * (size_t)i + 1];
            o_xyz[3 * (size_t)d + 2] = v_xyz[3 * (size_t)i + 2];
            o_key[d] = v_key[i];
            continue;
        }
        const uint32_t f = (rank > 0) ? x_lookup(low, cap, table, mask, v_key[i]) : NONE32;
        if (f == NONE32 || __ldcg(&ids_low[f]) == NONE32) {
            atomicAdd(flags + 5, 1u);
            gid[i] = NONE32;
        } else
            gid[i] = low_off + __ldcg(&ids_low[f]);
    }
}

// waits for every rank's record of this pass, then the prefixes (x_offsets_nb_kernel on the inbox)
__global__ void px_offsets_kernel(uint32_t* __restrict__ inbox, int world, uint32_t cap, uint32_t pass,
    uint32_t* __restrict__ voff, uint32_t* __restrict__ foff, unsigned* __restrict__ overflow,
    unsigned* __restrict__ need_ghost, unsigned* __restrict__ timeout)
{
    const uint32_t parity = pass & 1u;
    if ((int)threadIdx.x < world) px_wait(inbox + parity * 64 + 1 + threadIdx.x, pass, timeout);
    __syncthreads();
    if (threadIdx.x) return;
    uint32_t v = 0, f = 0, fv = 0, ft = 0;
    uint32_t* fvoff = foff + (world + 1);
    uint32_t* ftoff = fvoff + (world + 1);
    unsigned ovf = *overflow, ng = 0, bad = 0;
    for (int s = 0; s < world; ++s) {
        const uint32_t* r = px_record(inbox, cap, world, parity, s);
        voff[s] = v;
        foff[s] = f;
        fvoff[s] = fv;
        ftoff[s] = ft;
        v += __ldcg(r + 0);
        f += __ldcg(r + 1);
        fv += __ldcg(r + 3);
        ft += __ldcg(r + 4);
        ng |= __ldcg(r + 5);
        bad |= __ldcg(r + 6);
        if (__ldcg(r + 2) > cap) ovf = max(ovf, __ldcg(r + 2));
    }
    voff[world] = v;
    foff[world] = f;
    fvoff[world] = fv;
    ftoff[world] = ft;
    *overflow = ovf;
    *need_ghost = ng;
    need_ghost[1] = bad;
}

// ---- fused run + exchange (rin_run_exchange): the exchange is enqueued behind the run's kernels before the
// run's counts have reached the host, so every count is read from device memory ------------------------------
// header of this rank's count record from the device counters of the run
__global__ void fx_header_kernel(uint32_t* __restrict__ rec, const unsigned* __restrict__ n_faces,
    const unsigned* __restrict__ n_fv, const unsigned long long* __restrict__ n_zero, uint32_t known_degenerate,
    const unsigned* __restrict__ run_overflow, const unsigned* __restrict__ gen_err,
    const unsigned* __restrict__ gen_arena_overflow, const unsigned* __restrict__ n_bnd_faces)
{
    if (threadIdx.x || blockIdx.x) return;
    rec[1] = *n_faces;
    rec[3] = *n_fv;
    rec[4] = *n_faces; // interior iso-faces: one (tet, local face) pair each
    rec[5] = (known_degenerate || *n_zero != 0 || *n_bnd_faces != 0) ? 1u : 0u;
    rec[6] = (*run_overflow || *gen_err || *gen_arena_overflow) ? 1u : 0u;
}

// global vertex ids into the face vertex lists and the rebased face offsets, only when every rank's pass is
// complete (flags = {capacity overflow, -, unresolved vertices, degenerate, some run incomplete})
__global__ void __launch_bounds__(256) fx_apply_kernel(const unsigned* __restrict__ flags, uint32_t* __restrict__ f_verts,
    const uint32_t* __restrict__ gid, const unsigned* __restrict__ n_fv, uint32_t* __restrict__ f_off,
    uint32_t* __restrict__ f_toff, const unsigned* __restrict__ n_faces, const uint32_t* __restrict__ fv_off,
    const uint32_t* __restrict__ ft_off)
{
    if (flags[3] | flags[5] | flags[6] | flags[7]) return;
    const uint32_t nfv = *n_fv, nf1 = *n_faces + 1, a = *fv_off, b = *ft_off;
    const uint32_t step = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = i0; i < nfv; i += step) f_verts[i] = gid[f_verts[i]];
    for (uint32_t i = i0; i < nf1; i += step) {
        f_off[i] += a;
        f_toff[i] += b;
    }
}

} // namespace rin
