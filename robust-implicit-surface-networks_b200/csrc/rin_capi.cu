// C-ABI implementation (include/rin_b200.h): context, device buffers, stage orchestration.
// Host code here only sizes buffers and launches kernels; all per-vertex / per-tet work is in
// kernels_*.cuh.  There is no CPU fallback: without a CUDA device every entry point fails.
#include "../../include/rin_b200.h"
#include "kernels_mi.cuh"
#include "pipeline_ia.cuh"
#include "complexes.cuh"
#include "exchange.cuh"
#include "edges.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace rin;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

// stage events are recorded only while stage timing is on (rin_set_stage_timing); first / last always
#define EVREC(ev_)                                                                                 \
    do {                                                                                           \
        if (c->stage_timing || (ev_) == c->ev[0] || (ev_) == c->ev[ST_COUNT]) CK(cudaEventRecord((ev_), c->stream)); \
    } while (0)

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(RIN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

// grow-only device buffer
struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const
    {
        return static_cast<T*>(p);
    }
};

enum Stage {
    ST_EVAL = 0,   // function evaluation + signs          (load_functions + "func signs")
    ST_FILTER,     // active-function filter               ("filter")
    ST_CLASSIFY,   // table dispatch                       ("simp_arr(1 func)", "(2 func)")
    ST_GENERAL,    // general per-tet kernel               ("simp_arr(>=3 func)")
    ST_SCAN,       // output counts + offsets              ("extract mesh")
    ST_EMIT,       //                                      ("extract mesh")
    ST_DEDUP,      //                                      ("extract mesh")
    ST_VERTS,      // unique vertices + coordinates        ("compute xyz")
    ST_FACES,      //                                      ("extract mesh")
    ST_COUNT
};
const char* kStageNames[ST_COUNT] = {"eval+signs", "filter", "classify(lookup)", "general", "count+scan",
    "emit", "dedup", "verts+xyz", "faces"};


struct NcclApi
{
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, char[128], int) = nullptr; // ncclUniqueId is passed by value
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct UniqueId
{
    char internal[128];
};
NcclApi g_nccl;

struct Lut
{
    DevBuf lut1, lut2, blob;
    DevBuf cx2, lut2cx; // complete 2-plane complexes (start state of the general kernel) + key -> entry
    DevBuf cx3, lut3cx; // MI: complete 3-material complexes + key -> entry (uint32, 0xffffffff = none)
    uint32_t blob_bytes = 0;
    uint32_t n_keys3 = 0; // MI: realised keys of the 3-material table
    bool built = false;
    // host copies (tests / introspection)
    std::vector<uint16_t> h_lut1, h_lut2;
    std::vector<uint8_t> h_blob;
};

} // namespace

namespace {
struct Counters
{
    FilterCounters filt;
    GeneralCounters gen;
    ScanTotals scan;
    unsigned rank_tile, n_unique, n_bndry_faces, n_exact_classify, n_tie_faces;
    unsigned overflow; // OVF_* bits of the implicit-arrangement pass
    PassTotals tot;    // totals of the tile scan
    unsigned long long n_zero;
};
} // namespace

struct rin_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // fused run + exchange: the vertex exchange runs beside the face kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    float x_chain_ms = 0; // device time of the forked exchange chain of the last fused pass
    cudaEvent_t ev_x[10] = {}; // per-kernel brackets of the peer exchange chain (stage timing on)
    float x_part_ms[10] = {};
    int x_parts = 0;
    // peer-memory exchange (exchange.cuh): this rank's inbox and the peers' inboxes mapped through CUDA IPC
    uint32_t* p_inbox = nullptr;
    uint32_t* p_peer[PX_MAX_WORLD] = {};
    uint32_t p_cap = 0, p_pass = 0;
    bool p_ready = false, p_failed = false;
    DevBuf p_stage;
    int sm_count = 148;
    size_t smem_per_sm = 227 * 1024;
    // mesh
    uint64_t V = 0, T = 0;
    uint32_t VS = 0;      // row stride of vals / vmask: V rounded up to 16 entries (128-byte rows)
    uint32_t grid_R = 0;  // != 0: the mesh is generate_tet_mesh(grid_R) (structured filter, axis tables)
    DevBuf pts, tets, axes;
    uint64_t t_first = 0, t_count = 0; // tets processed by a run (own range + ghosts)
    uint64_t own_first = 0, own_count = 0, ghost_lo = 0, ghost_hi = 0; // rin_set_ghost_tets
    uint64_t res_first = 0, res_count = 0; // tets resident on the device (a slice after rin_set_mesh_host_range)
    uint32_t ghost_v_lo = 0; // local vertices [0, ghost_v_lo) of the last run were first created by a ghost below
    DevBuf g_off, g_verts, g_toff, g_tets, g_funcs; // face arrays without the faces created by ghosts
    // functions / values
    uint32_t F = 0;
    DevBuf funcs;
    bool have_funcs = false, have_values = false;
    DevBuf rowmajor; // caller-provided values (device copy)
    DevBuf vals, vmask, vmask16;
    // work buffers
    DevBuf counters; // small zeroed block: FilterCounters | GeneralCounters | ScanTotals | misc
    DevBuf status;   // look-back status words
    DevBuf tl_tet, tl_mask, tile_cnt, tile_off; // tile-local filter output
    DevBuf tl_ref, tile_tot, tile_pre;          // implicit-arrangement pass: record refs, tile totals / prefixes
    // sizes learnt from the previous pass (0 = unknown: the next pass sizes its buffers after the tile scan)
    uint32_t h_act = 0, h_list = 0, h_cand = 0, h_face = 0, h_fv = 0, h_unique = 0;
    uint32_t h_act_mi = 0, h_gen_mi = 0; // material-interface pass: active / general tets of the previous pass (+ margin)
    uint32_t table_size = 0; // vertex hash table slots of the last IA pass
    DevBuf act_tet, act_mask, rec_ref, general_list, big_list, arena, offs;
    DevBuf cand_key, cand_pay, cand_src, face_hdr, fv_ref;
    DevBuf table, slot_of, rep, vid;
    DevBuf tmp_fverts, bfkeys, frep, fdup, fpos, bf_mask; // degenerate boundary-face dedup
    DevBuf m_cnt, m_off, m_vmap, m_fmap, fpartner;        // cell-grouping maps (rin_tet_maps)
    bool ia_bndry_faces = false;                          // last run took the boundary-face path
    uint32_t attr_set = 0; // kernel attributes already set on this context's device (bit W: mid tier, bit 16: eval_mi)
    bool skip_mid = false; // the last IA pass had no tet for the mid tier: its (empty) launch is left out, see run_ia_w
    uint64_t m_nv = 0, m_nf = 0;
    bool maps_ready = false;
    // outputs
    DevBuf v_tet, v_local, v_size, v_simplex, v_funcs, v_xyz, v_key;
    // sharded runs
    DevBuf o_tet, o_local, o_size, o_simplex, o_funcs, o_xyz, o_key, own_flag, own_idx, gid, fkeys, fgids, ftable,
        bkeys, bids;
    bool marked = false, finalized = false;
    // NCCL exchange
    void* nccl_comm = nullptr;
    int x_rank = 0, x_world = 1;
    uint32_t x_cap = 0, x_lo = 1, x_hi = 0;
    bool x_window = false;
    bool x_neighbours = false;          // every rank shares vertices with ranks r-1 / r+1 only (slab sharding)
    uint32_t x_up_lo = 1, x_up_hi = 0;  // vertex window shared with rank r+1 (empty: lo > hi)
    bool x_has_low = false, x_has_up = false;
    bool x_fusable = false;   // neighbour protocol and no empty rank: rin_run_exchange may fuse
    bool fuse = false;        // the current rin_run enqueues the exchange behind its kernels
    bool fuse_disabled = false; // degenerate inputs: the ranks agreed to use the two-call path
    uint32_t* h_xsmall = nullptr; // pinned mirror of x_small
    DevBuf x_ids_up, x_ids_low, x_cnt;
    uint64_t x_offsets[8] = {}; // offset / total of vertices, faces, face-vertex entries, face-tet pairs (last exchange)
    DevBuf x_send, x_recv1, x_recv2, x_table, x_small, x_status;
    DevBuf cx_out; // rin_get_complexes output arena
    DevBuf e_key, e_slot, e_table, e_verts, e_of_face, e_cnt, e_off, e_pairs; // rin_mesh_edges
    uint64_t n_edges = 0;
    bool edges_ready = false;
    void* h_pinned = nullptr; // pinned host mirror of the device counters (cheap read-back)
    uint32_t n_local_verts = 0, n_own = 0;
    DevBuf f_off, f_verts, f_toff, f_tets, f_funcs;
    uint32_t act_cap = 0;
    Lut lut_ia, lut_mi;
    rin_counts counts{};
    int last_mode = -1;
    uint32_t last_flags = 0;
    float stage_ms[ST_COUNT] = {};
    cudaEvent_t ev[ST_COUNT + 1] = {};
    cudaEvent_t kev[4] = {}; // tight brackets around the eval and filter kernels
    float kernel_ms[2] = {};
    float total_ms = 0;
    uint32_t v_first = 0, v_count = 0; // vertex range touched by the tet range
    bool ran = false;
    // snapshot of the inputs of the last successful run (what the download / introspection calls refer to)
    uint32_t run_F = 0, run_VS = 0;
    uint64_t run_V = 0, run_t_first = 0, run_t_count = 0;
    int launches = 0; // kernels launched by the last rin_run
    bool stage_timing = true;
};

namespace {


// every change of the inputs invalidates the results of the previous run (downloads then fail with
// RIN_ERR_STATE instead of reading buffers sized for other inputs)
void invalidate(rin_ctx* c)
{
    c->ran = false;
    c->maps_ready = false;
    c->edges_ready = false;
    c->marked = c->finalized = false;
}

int grid_for(uint64_t n, int threads, int sm_count, int per_sm = 8)
{
    uint64_t b = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)sm_count * per_sm;
    return (int)std::max<uint64_t>(1, std::min(b, cap));
}

uint32_t words_for(uint32_t F)
{
    return (F + 31) / 32;
}

template <int W>
int run_ia_w(rin_ctx* c, uint32_t flags);
template <int W>
int run_mi_w(rin_ctx* c, uint32_t flags);

int build_ia_tables(rin_ctx* c);
int build_mi_tables(rin_ctx* c);

} // namespace

extern "C" {

const char* rin_last_error(void)
{
    return g_err.c_str();
}

int rin_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int rin_create(int device, rin_ctx** out)
{
    if (!out) return fail(RIN_ERR_ARG, "rin_create: null out");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return fail(RIN_ERR_NO_DEVICE, "rin_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(RIN_ERR_ARG, "rin_create: bad device index");
    CK(cudaSetDevice(device));
    auto* c = new rin_ctx;
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    for (auto& ev : c->ev)
        if (e == cudaSuccess) e = cudaEventCreate(&ev);
    for (auto& ev : c->kev)
        if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_pinned, sizeof(Counters), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        rin_destroy(c); // frees whatever was created
        return fail(RIN_ERR_CUDA, std::string("rin_create: ") + cudaGetErrorString(e));
    }
    c->sm_count = prop.multiProcessorCount;
    c->smem_per_sm = prop.sharedMemPerMultiprocessor;
    *out = c;
    return RIN_OK;
}

extern "C++" {
namespace {
void release_peers(rin_ctx* c)
{
    for (int s = 0; s < PX_MAX_WORLD; ++s) {
        if (c->p_peer[s] && c->p_peer[s] != c->p_inbox) cudaIpcCloseMemHandle(c->p_peer[s]);
        c->p_peer[s] = nullptr;
    }
    if (c->p_inbox) cudaFree(c->p_inbox);
    c->p_inbox = nullptr;
    c->p_ready = false;
    c->p_cap = 0;
}
} // namespace
}

void rin_destroy(rin_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->pts, &c->tets, &c->funcs, &c->rowmajor, &c->vals, &c->vmask, &c->vmask16, &c->counters,
        &c->status, &c->tl_tet, &c->tl_mask, &c->tile_cnt, &c->tile_off, &c->tl_ref, &c->tile_tot, &c->tile_pre, &c->axes, &c->act_tet, &c->act_mask, &c->rec_ref, &c->general_list, &c->big_list, &c->arena, &c->offs, &c->m_cnt, &c->m_off, &c->m_vmap, &c->m_fmap, &c->fpartner,
        &c->cand_key, &c->cand_pay, &c->cand_src, &c->face_hdr, &c->fv_ref, &c->table, &c->slot_of, &c->rep, &c->vid, &c->tmp_fverts, &c->bfkeys, &c->frep, &c->fdup, &c->fpos, &c->bf_mask,
        &c->v_tet, &c->v_local, &c->v_size, &c->v_simplex, &c->v_funcs, &c->v_xyz, &c->v_key, &c->o_tet, &c->o_local,
        &c->o_size, &c->o_simplex, &c->o_funcs, &c->o_xyz, &c->o_key, &c->own_flag, &c->own_idx, &c->gid, &c->fkeys,
        &c->fgids, &c->ftable, &c->bkeys, &c->bids, &c->x_send, &c->x_recv1, &c->x_recv2, &c->x_table, &c->x_small, &c->x_ids_up, &c->x_ids_low, &c->x_cnt, &c->e_key, &c->e_slot, &c->e_table, &c->e_verts, &c->e_of_face,
        &c->e_cnt, &c->e_off, &c->e_pairs, &c->cx_out, &c->f_off, &c->f_verts,
        &c->f_toff, &c->f_tets, &c->f_funcs, &c->lut_ia.lut1, &c->lut_ia.lut2, &c->lut_ia.blob, &c->lut_ia.cx2, &c->lut_ia.lut2cx,
        &c->lut_mi.lut1, &c->lut_mi.lut2, &c->lut_mi.blob, &c->lut_mi.cx3, &c->lut_mi.lut3cx, &c->g_off, &c->g_verts,
        &c->g_toff, &c->g_tets, &c->g_funcs, &c->x_status, &c->p_stage};
    for (auto* b : bufs) b->release();
    for (auto& e : c->ev_x)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->kev)
        if (e) cudaEventDestroy(e);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_xsmall) cudaFreeHost(c->h_xsmall);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    release_peers(c);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
}

int rin_set_mesh_host(rin_ctx* c, const double* pts, uint64_t n_pts, const void* tets, uint64_t n_tets,
    int index_bytes)
{
    if (!c || !pts || !tets) return fail(RIN_ERR_ARG, "rin_set_mesh_host: null argument");
    if (index_bytes != 4 && index_bytes != 8) return fail(RIN_ERR_ARG, "index_bytes must be 4 or 8");
    if (n_pts >= 0xffffffffull || n_tets >= 0x7fffffffull)
        return fail(RIN_ERR_ARG, "mesh too large for 32-bit device indices");
    CK(cudaSetDevice(c->device));
    CK(c->pts.ensure(n_pts * 24));
    CK(c->tets.ensure(n_tets * 16));
    CK(cudaMemcpyAsync(c->pts.p, pts, n_pts * 24, cudaMemcpyHostToDevice, c->stream));
    if (index_bytes == 4) {
        CK(cudaMemcpyAsync(c->tets.p, tets, n_tets * 16, cudaMemcpyHostToDevice, c->stream));
    } else {
        // the reference hands std::array<size_t,4>: upload raw, narrow on the device
        CK(c->rowmajor.ensure(n_tets * 32));
        CK(cudaMemcpyAsync(c->rowmajor.p, tets, n_tets * 32, cudaMemcpyHostToDevice, c->stream));
        narrow_tets_kernel<<<grid_for(n_tets, 256, c->sm_count), 256, 0, c->stream>>>(
            c->rowmajor.as<uint64_t>(), n_tets, c->tets.as<uint4>());
        CK(cudaGetLastError());
    }
    const bool same_shape = c->V == n_pts && c->T == n_tets && c->t_first == 0 && c->t_count == n_tets && c->grid_R == 0;
    c->V = n_pts;
    c->VS = (uint32_t)((n_pts + 15) & ~15ull);
    c->grid_R = 0;
    c->T = n_tets;
    c->t_first = c->own_first = c->res_first = 0;
    c->t_count = c->own_count = c->res_count = n_tets;
    c->ghost_lo = c->ghost_hi = 0;
    c->v_first = c->v_count = 0;
    c->have_values = false;
    if (!same_shape) { // sizes learnt from the previous pass stay valid hints for a mesh of the same shape
        c->x_window = false;
        c->h_act_mi = c->h_act = c->h_list = c->h_cand = c->h_face = c->h_fv = c->h_unique = 0;
    }
    invalidate(c);
    return RIN_OK;
}

// One rank's slice of a mesh that is sharded by contiguous tet ranges: device buffers are allocated for the
// whole mesh (indices stay global, so keys agree between ranks) but only the slice crosses PCIe.
int rin_set_mesh_host_range(rin_ctx* c, uint64_t n_pts, uint64_t n_tets, const double* pts_slice, uint64_t v_first,
    uint64_t v_count, const void* tets_slice, uint64_t t_first, uint64_t t_count, int index_bytes)
{
    if (!c || (v_count && !pts_slice) || (t_count && !tets_slice)) return fail(RIN_ERR_ARG, "rin_set_mesh_host_range: null argument");
    if (index_bytes != 4 && index_bytes != 8) return fail(RIN_ERR_ARG, "index_bytes must be 4 or 8");
    if (n_pts >= 0xffffffffull || n_tets >= 0x7fffffffull) return fail(RIN_ERR_ARG, "mesh too large for 32-bit device indices");
    if (v_first > n_pts || v_count > n_pts - v_first || t_first > n_tets || t_count > n_tets - t_first)
        return fail(RIN_ERR_ARG, "rin_set_mesh_host_range: slice out of bounds");
    CK(cudaSetDevice(c->device));
    CK(c->pts.ensure(n_pts * 24));
    CK(c->tets.ensure(n_tets * 16));
    if (v_count)
        CK(cudaMemcpyAsync(c->pts.as<double>() + 3 * v_first, pts_slice, v_count * 24, cudaMemcpyHostToDevice, c->stream));
    if (t_count) {
        if (index_bytes == 4) {
            CK(cudaMemcpyAsync(c->tets.as<uint4>() + t_first, tets_slice, t_count * 16, cudaMemcpyHostToDevice, c->stream));
        } else {
            CK(c->rowmajor.ensure(t_count * 32));
            CK(cudaMemcpyAsync(c->rowmajor.p, tets_slice, t_count * 32, cudaMemcpyHostToDevice, c->stream));
            narrow_tets_kernel<<<grid_for(t_count, 256, c->sm_count), 256, 0, c->stream>>>(
                c->rowmajor.as<uint64_t>(), t_count, c->tets.as<uint4>() + t_first);
            CK(cudaGetLastError());
        }
    }
    const bool same_shape = c->V == n_pts && c->T == n_tets && c->t_first == t_first && c->t_count == t_count &&
                            c->grid_R == 0;
    c->V = n_pts;
    c->VS = (uint32_t)((n_pts + 15) & ~15ull);
    c->grid_R = 0;
    c->T = n_tets;
    c->t_first = c->own_first = c->res_first = t_first;
    c->t_count = c->own_count = c->res_count = t_count;
    c->ghost_lo = c->ghost_hi = 0;
    c->v_first = (uint32_t)v_first;
    c->v_count = (uint32_t)v_count;
    c->have_values = false;
    if (!same_shape) {
        c->x_window = false;
        c->h_act_mi = c->h_act = c->h_list = c->h_cand = c->h_face = c->h_fv = c->h_unique = 0;
    }
    invalidate(c);
    return RIN_OK;
}

int rin_set_values_host_range(rin_ctx* c, const double* vals_slice, uint64_t v_first, uint64_t v_count, uint32_t F)
{
    if (!c || (v_count && !vals_slice) || F == 0) return fail(RIN_ERR_ARG, "rin_set_values_host_range: bad argument");
    if (v_first > c->V || v_count > c->V - v_first) return fail(RIN_ERR_ARG, "rin_set_values_host_range: slice out of bounds");
    if (F > RIN_MAX_FUNCS) return fail(RIN_ERR_ARG, "more than 128 functions are not supported by this build");
    CK(cudaSetDevice(c->device));
    CK(c->rowmajor.ensure(c->V * F * 8));
    if (v_count)
        CK(cudaMemcpyAsync(c->rowmajor.as<double>() + v_first * F, vals_slice, v_count * F * 8, cudaMemcpyHostToDevice,
            c->stream));
    c->F = F;
    c->have_values = true;
    c->have_funcs = false;
    invalidate(c);
    return RIN_OK;
}

int rin_generate_grid(rin_ctx* c, uint32_t R, const double bmin[3], const double bmax[3])
{
    if (!c || R == 0) return fail(RIN_ERR_ARG, "rin_generate_grid: bad argument");
    const uint64_t N = (uint64_t)R + 1;
    const uint64_t V = N * N * N, T = (uint64_t)R * R * R * 5;
    if (V >= 0xffffffffull || T >= 0x7fffffffull) return fail(RIN_ERR_ARG, "grid too large");
    CK(cudaSetDevice(c->device));
    CK(c->pts.ensure(V * 24));
    CK(c->tets.ensure(T * 16));
    grid_points_kernel<<<grid_for(V, 256, c->sm_count), 256, 0, c->stream>>>((uint32_t)N,
        make_double3(bmin[0], bmin[1], bmin[2]), make_double3(bmax[0], bmax[1], bmax[2]), c->pts.as<double>());
    grid_tets_kernel<<<grid_for(T, 256, c->sm_count), 256, 0, c->stream>>>(R, c->tets.as<uint4>());
    // axis tables: the evaluation kernel of a generated grid reads these instead of the points
    CK(c->axes.ensure(3 * N * 8));
    grid_axes_kernel<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>((uint32_t)N,
        make_double3(bmin[0], bmin[1], bmin[2]), make_double3(bmax[0], bmax[1], bmax[2]), c->axes.as<double>());
    CK(cudaGetLastError());
    c->V = V;
    c->VS = (uint32_t)((V + 15) & ~15ull);
    c->grid_R = R;
    c->T = T;
    c->t_first = c->own_first = c->res_first = 0;
    c->t_count = c->own_count = c->res_count = T;
    c->ghost_lo = c->ghost_hi = 0;
    c->v_first = c->v_count = 0;
    c->have_values = false;
    c->h_act_mi = c->h_act = c->h_list = c->h_cand = c->h_face = c->h_fv = c->h_unique = 0;
    invalidate(c);
    return RIN_OK;
}

namespace {
// tets processed = own range widened by the ghosts; only the vertices they reference are evaluated
int apply_tet_range(rin_ctx* c)
{
    const uint64_t lo = std::min<uint64_t>(c->ghost_lo, c->own_first);
    const uint64_t hi = std::min<uint64_t>(c->ghost_hi, c->T - (c->own_first + c->own_count));
    if (c->own_count && (c->own_first - lo < c->res_first || c->own_first + c->own_count + hi > c->res_first + c->res_count))
        return fail(RIN_ERR_STATE, "tet range (with ghosts) outside the slice uploaded by rin_set_mesh_host_range");
    c->t_first = c->own_first - lo;
    c->t_count = c->own_count ? c->own_count + lo + hi : 0; // an empty rank stays empty
    c->v_first = 0;
    c->v_count = 0;
    c->x_window = false;
    c->h_act_mi = c->h_act = c->h_list = c->h_cand = c->h_face = c->h_fv = c->h_unique = 0;
    invalidate(c);
    if (c->t_count != 0 && (c->t_first != 0 || c->t_count != c->T)) {
        CK(cudaSetDevice(c->device));
        CK(c->counters.ensure(sizeof(Counters)));
        uint32_t init[2] = {0xffffffffu, 0u};
        uint32_t* d = c->counters.as<uint32_t>();
        CK(cudaMemcpyAsync(d, init, 8, cudaMemcpyHostToDevice, c->stream));
        vertex_range_kernel<<<grid_for(c->t_count, 256, c->sm_count), 256, 0, c->stream>>>(c->tets.as<uint4>(),
            (uint32_t)c->t_first, (uint32_t)c->t_count, d);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(init, d, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->v_first = init[0];
        c->v_count = init[1] - init[0] + 1;
    }
    return RIN_OK;
}
} // namespace

int rin_set_tet_range(rin_ctx* c, uint64_t first, uint64_t count)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (first > c->T) return fail(RIN_ERR_ARG, "tet range out of bounds");
    if (count == RIN_TET_RANGE_ALL) count = c->T - first;
    if (count > c->T - first) return fail(RIN_ERR_ARG, "tet range out of bounds");
    c->own_first = first;
    c->own_count = count; // 0: an empty range, rin_run then returns an empty result
    c->ghost_lo = c->ghost_hi = 0;
    return apply_tet_range(c);
}

int rin_set_ghost_tets(rin_ctx* c, uint64_t below, uint64_t above)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (c->T == 0) return fail(RIN_ERR_STATE, "rin_set_ghost_tets: no mesh");
    c->ghost_lo = below;
    c->ghost_hi = above;
    return apply_tet_range(c);
}

int rin_set_functions(rin_ctx* c, const rin_func_desc* funcs, uint32_t F)
{
    if (!c || !funcs || F == 0) return fail(RIN_ERR_ARG, "rin_set_functions: bad argument");
    if (F > RIN_MAX_FUNCS) return fail(RIN_ERR_ARG, "more than 128 functions are not supported by this build");
    CK(cudaSetDevice(c->device));
    CK(c->funcs.ensure(F * sizeof(rin_func_desc)));
    CK(cudaMemcpyAsync(c->funcs.p, funcs, F * sizeof(rin_func_desc), cudaMemcpyHostToDevice, c->stream));
    c->F = F;
    c->have_funcs = true;
    c->have_values = false;
    invalidate(c);
    return RIN_OK;
}

int rin_set_values_host(rin_ctx* c, const double* vals, uint64_t n_pts, uint32_t F)
{
    if (!c || !vals || F == 0) return fail(RIN_ERR_ARG, "rin_set_values_host: bad argument");
    if (n_pts != c->V) return fail(RIN_ERR_ARG, "rin_set_values_host: row count != number of points");
    if (F > RIN_MAX_FUNCS) return fail(RIN_ERR_ARG, "more than 128 functions are not supported by this build");
    CK(cudaSetDevice(c->device));
    CK(c->rowmajor.ensure(n_pts * F * 8));
    CK(cudaMemcpyAsync(c->rowmajor.p, vals, n_pts * F * 8, cudaMemcpyHostToDevice, c->stream));
    c->F = F;
    c->have_values = true;
    c->have_funcs = false;
    invalidate(c);
    return RIN_OK;
}

namespace {
// rin_set_ghost_tets: keep what the rank's own tets created (exchange.cuh, GhostBounds)
int trim_ghosts(rin_ctx* c)
{
    c->ghost_v_lo = 0;
    if (c->t_first == c->own_first && c->t_count == c->own_count) return RIN_OK;
    cudaStream_t s = c->stream;
    const uint32_t NV = c->n_local_verts, NF = (uint32_t)c->counts.num_faces;
    CK(c->x_small.ensure(4096 + 64 * (size_t)std::max(c->x_world, 1)));
    GhostBounds* d_b = reinterpret_cast<GhostBounds*>(c->x_small.as<uint32_t>() + 512);
    ghost_bounds_kernel<<<1, 1, 0, s>>>(c->v_tet.as<uint32_t>(), NV, c->f_off.as<uint32_t>(), c->f_toff.as<uint32_t>(),
        c->f_tets.as<uint32_t>(), NF, (uint32_t)c->own_first, (uint32_t)(c->own_first + c->own_count), d_b);
    CK(cudaGetLastError());
    GhostBounds b;
    CK(cudaMemcpyAsync(&b, d_b, sizeof(b), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint32_t nf = b.f_hi - b.f_lo, nfv = b.fv_hi - b.fv_lo, nft = b.ft_hi - b.ft_lo;
    CK(c->g_off.ensure(((size_t)nf + 1) * 4));
    CK(c->g_toff.ensure(((size_t)nf + 1) * 4));
    CK(c->g_funcs.ensure(std::max<size_t>(nf, 1) * 8));
    CK(c->g_verts.ensure(std::max<size_t>(nfv, 1) * 4));
    CK(c->g_tets.ensure(std::max<size_t>(nft, 1) * 8));
    ghost_slice_faces_kernel<<<grid_for(std::max<uint64_t>({nf + 1ull, nfv, 2ull * nft}), 256, c->sm_count), 256, 0, s>>>(b,
        c->f_off.as<uint32_t>(), c->f_verts.as<uint32_t>(), c->f_toff.as<uint32_t>(), c->f_tets.as<uint32_t>(),
        c->f_funcs.as<uint32_t>(), c->g_off.as<uint32_t>(), c->g_verts.as<uint32_t>(), c->g_toff.as<uint32_t>(),
        c->g_tets.as<uint32_t>(), c->g_funcs.as<uint32_t>());
    CK(cudaGetLastError());
    std::swap(c->f_off, c->g_off);
    std::swap(c->f_verts, c->g_verts);
    std::swap(c->f_toff, c->g_toff);
    std::swap(c->f_tets, c->g_tets);
    std::swap(c->f_funcs, c->g_funcs);
    c->ghost_v_lo = b.v_lo;
    c->n_local_verts = c->n_own = b.v_hi; // what the ghosts above created is gone
    c->counts.num_verts = b.v_hi;
    c->counts.num_faces = nf;
    c->counts.num_face_verts = nfv;
    c->counts.num_face_tets = nft;
    return RIN_OK;
}
} // namespace

int rin_run(rin_ctx* c, int mode, uint32_t flags)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (c->T == 0 || c->V == 0) return fail(RIN_ERR_STATE, "rin_run: no mesh");
    if (!c->have_funcs && !c->have_values) return fail(RIN_ERR_STATE, "rin_run: no functions / values");
    CK(cudaSetDevice(c->device));
    invalidate(c); // a failed run must not leave the previous run's results readable
    auto snapshot = [&]() {
        c->last_mode = mode;
        c->last_flags = flags;
        c->run_F = c->F;
        c->run_V = c->V;
        c->run_VS = c->VS;
        c->run_t_first = c->t_first;
        c->run_t_count = c->t_count;
        c->ran = true;
    };
    if (!(flags & RIN_FLAG_USE_LOOKUP)) flags &= ~RIN_FLAG_USE_SECONDARY_LOOKUP; // :38-40
    if (mode == RIN_MODE_IA) {
        if ((flags & RIN_FLAG_USE_LOOKUP) && !c->lut_ia.built) {
            int rc = build_ia_tables(c);
            if (rc) return rc;
        }
        int rc;
        switch (words_for(c->F)) {
        case 1: rc = run_ia_w<1>(c, flags); break;
        case 2: rc = run_ia_w<2>(c, flags); break;
        case 3: rc = run_ia_w<3>(c, flags); break;
        case 4: rc = run_ia_w<4>(c, flags); break;
        default: return fail(RIN_ERR_ARG, "more than 128 functions are not supported by this build");
        }
        if (rc == RIN_OK) rc = trim_ghosts(c);
        if (rc == RIN_OK) snapshot();
        return rc;
    }
    if (mode == RIN_MODE_MI) {
        if ((flags & RIN_FLAG_USE_LOOKUP) && !c->lut_mi.built) {
            int rc = build_mi_tables(c);
            if (rc) return rc;
        }
        int rc;
        switch (words_for(c->F)) {
        case 1: rc = run_mi_w<1>(c, flags); break;
        case 2: rc = run_mi_w<2>(c, flags); break;
        case 3: rc = run_mi_w<3>(c, flags); break;
        case 4: rc = run_mi_w<4>(c, flags); break;
        default: return fail(RIN_ERR_ARG, "more than 128 materials are not supported by this build");
        }
        if (rc == RIN_OK) rc = trim_ghosts(c);
        if (rc == RIN_OK) snapshot();
        return rc;
    }
    return fail(RIN_ERR_ARG, "rin_run: unknown mode");
}

int rin_get_counts(const rin_ctx* c, rin_counts* out)
{
    if (!c || !out) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran) return fail(RIN_ERR_STATE, "no finished run");
    *out = c->counts;
    return RIN_OK;
}

int rin_download_mesh(rin_ctx* c, rin_mesh_out* o)
{
    if (!c || !o) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran) return fail(RIN_ERR_STATE, "no finished run");
    CK(cudaSetDevice(c->device));
    const rin_counts& n = c->counts;
    auto dl = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t {
        if (!dst || bytes == 0) return cudaSuccess;
        return cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    CK(dl(o->vert_tet, c->v_tet, n.num_verts * 4));
    CK(dl(o->vert_local, c->v_local, n.num_verts));
    CK(dl(o->vert_simplex_size, c->v_size, n.num_verts));
    CK(dl(o->vert_simplex, c->v_simplex, n.num_verts * 16));
    CK(dl(o->vert_funcs, c->v_funcs, n.num_verts * 16));
    CK(dl(o->vert_xyz, c->v_xyz, n.num_verts * 24));
    CK(dl(o->face_offsets, c->f_off, (n.num_faces + 1) * 4));
    CK(dl(o->face_verts, c->f_verts, n.num_face_verts * 4));
    CK(dl(o->face_tet_offsets, c->f_toff, (n.num_faces + 1) * 4));
    CK(dl(o->face_tets, c->f_tets, n.num_face_tets * 8));
    CK(dl(o->face_funcs, c->f_funcs, n.num_faces * 8));
    CK(cudaStreamSynchronize(c->stream));
    return RIN_OK;
}

int rin_download_active(rin_ctx* c, uint32_t* func_in_tet, uint64_t* start)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (!c->ran) return fail(RIN_ERR_STATE, "no finished run");
    CK(cudaSetDevice(c->device));
    const uint32_t A = (uint32_t)c->counts.num_intersecting_tet, W = words_for(c->run_F);
    std::vector<uint32_t> at(A), am((size_t)A * W);
    if (A) {
        CK(cudaMemcpyAsync(at.data(), c->act_tet.p, (size_t)A * 4, cudaMemcpyDeviceToHost, c->stream));
        for (uint32_t w = 0; w < W; ++w)
            CK(cudaMemcpyAsync(am.data() + (size_t)w * A, c->act_mask.as<uint32_t>() + (size_t)w * c->act_cap,
                (size_t)A * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    // format conversion only: expand the compact (tet, mask) list into the reference's CRS arrays
    uint64_t pos = 0, a = 0;
    for (uint64_t t = 0; t < c->run_t_count; ++t) {
        if (start) start[t] = pos;
        if (a < A && at[a] == c->run_t_first + t) {
            for (uint32_t w = 0; w < W; ++w) {
                uint32_t m = am[(size_t)w * A + a];
                while (m) {
                    int b = __builtin_ctz(m);
                    m &= m - 1;
                    if (func_in_tet) func_in_tet[pos] = w * 32 + b;
                    ++pos;
                }
            }
            ++a;
        }
    }
    if (start) start[c->run_t_count] = pos;
    return RIN_OK;
}

int rin_download_active_tets(rin_ctx* c, uint32_t* out)
{
    if (!c || !out) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran) return fail(RIN_ERR_STATE, "no finished run");
    CK(cudaSetDevice(c->device));
    const size_t A = (size_t)c->counts.num_intersecting_tet;
    if (A) CK(cudaMemcpyAsync(out, c->act_tet.p, A * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RIN_OK;
}

int rin_download_values(rin_ctx* c, double* out)
{
    if (!c || !out) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran) return fail(RIN_ERR_STATE, "no finished run");
    CK(cudaSetDevice(c->device));
    const uint64_t V = c->run_V, VS = c->run_VS;
    const uint32_t F = c->run_F;
    // rows of V values at a pitch of VS (the padding between rows is never written)
    std::vector<double> soa((size_t)V * F);
    CK(cudaMemcpy2DAsync(soa.data(), V * 8, c->vals.p, VS * 8, V * 8, F, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (uint64_t v = 0; v < V; ++v)
        for (uint32_t f = 0; f < F; ++f) out[v * F + f] = soa[(size_t)f * V + v];
    return RIN_OK;
}

int rin_download_grid(rin_ctx* c, double* pts, uint32_t* tets)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    if (pts) CK(cudaMemcpyAsync(pts, c->pts.p, c->V * 24, cudaMemcpyDeviceToHost, c->stream));
    if (tets) CK(cudaMemcpyAsync(tets, c->tets.p, c->T * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RIN_OK;
}

int rin_set_stage_timing(rin_ctx* c, int on)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    c->stage_timing = on != 0;
    return RIN_OK;
}

int rin_get_launch_count(const rin_ctx* c)
{
    return c ? c->launches : 0;
}

int rin_get_stage_times(const rin_ctx* c, float* ms, int capacity)
{
    if (!c || !ms) return fail(RIN_ERR_ARG, "null argument");
    for (int i = 0; i < capacity && i < ST_COUNT; ++i) ms[i] = c->stage_ms[i];
    return RIN_OK;
}
int rin_get_kernel_times(const rin_ctx* c, float* eval_ms, float* filter_ms, float* total_ms)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (eval_ms) *eval_ms = c->kernel_ms[0];
    if (filter_ms) *filter_ms = c->kernel_ms[1];
    if (total_ms) *total_ms = c->total_ms;
    return RIN_OK;
}
const char* rin_stage_name(int i)
{
    return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : "";
}
int rin_num_stages(void)
{
    return ST_COUNT;
}

int rin_run_host(rin_ctx* c, int mode, uint32_t flags, const double* pts, uint64_t n_pts, const void* tets,
    uint64_t n_tets, int index_bytes, const double* vals, uint32_t F, rin_counts* counts)
{
    int rc = rin_set_mesh_host(c, pts, n_pts, tets, n_tets, index_bytes);
    if (rc) return rc;
    rc = rin_set_values_host(c, vals, n_pts, F);
    if (rc) return rc;
    rc = rin_run(c, mode, flags);
    if (rc) return rc;
    if (counts) *counts = c->counts;
    return RIN_OK;
}

extern "C++" {
namespace {
template <int W>
int get_complexes_w(rin_ctx* c, int mode, const uint32_t* d_req, uint32_t n, uint32_t* d_out, uint32_t out_cap,
    uint2* d_span, ComplexCounters* d_cc)
{
    const uint32_t A = (uint32_t)c->counts.num_intersecting_tet;
    if (mode == RIN_MODE_IA)
        complexes_ia_kernel<W><<<grid_for(n, GEN_THREADS, c->sm_count, 4), GEN_THREADS, 0, c->stream>>>(
            c->tets.as<uint4>(), c->act_tet.as<uint32_t>(), c->act_mask.as<uint32_t>(), c->act_cap, A,
            c->vals.as<double>(), c->run_VS, d_req, n, d_out, out_cap, d_span, d_cc);
    else
        complexes_mi_kernel<W><<<grid_for(n, GEN_THREADS, c->sm_count, 4), GEN_THREADS, 0, c->stream>>>(
            c->tets.as<uint4>(), c->act_tet.as<uint32_t>(), c->act_mask.as<uint32_t>(), c->act_cap, A,
            c->vals.as<double>(), c->run_VS, d_req, n, d_out, out_cap, d_span, d_cc);
    CK(cudaGetLastError());
    return RIN_OK;
}
} // namespace
} // extern "C++"

// flags are unused: the general algorithm and the tables give identical complexes
int rin_get_complexes(rin_ctx* c, int mode, uint32_t, const uint64_t* tet_ids, uint64_t n, uint64_t* offsets,
    uint32_t* words, uint64_t* n_words)
{
    if (!c || (n && !tet_ids) || !n_words) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran || c->last_mode != mode) return fail(RIN_ERR_STATE, "rin_get_complexes: no finished run of that mode");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    std::vector<uint32_t> req(n);
    for (uint64_t i = 0; i < n; ++i) {
        if (tet_ids[i] >= c->T) return fail(RIN_ERR_ARG, "rin_get_complexes: tet id out of range");
        req[i] = (uint32_t)tet_ids[i];
    }
    DevBuf d_req, d_span, d_cc;
    auto cleanup = [&]() {
        d_req.release();
        d_span.release();
        d_cc.release();
    };
    cudaError_t e;
    if ((e = d_req.ensure(std::max<uint64_t>(n, 1) * 4)) != cudaSuccess || (e = d_span.ensure(std::max<uint64_t>(n, 1) * 8)) != cudaSuccess ||
        (e = d_cc.ensure(sizeof(ComplexCounters))) != cudaSuccess) {
        cleanup();
        return fail(RIN_ERR_CUDA, cudaGetErrorString(e));
    }
    if (n) cudaMemcpyAsync(d_req.p, req.data(), n * 4, cudaMemcpyHostToDevice, s);
    std::vector<uint2> span(n);
    ComplexCounters hc{};
    size_t cap_words = std::max<size_t>(c->cx_out.cap / 4, 1 << 18);
    for (int attempt = 0;; ++attempt) {
        if ((e = c->cx_out.ensure(cap_words * 4)) != cudaSuccess) {
            cleanup();
            return fail(RIN_ERR_CUDA, cudaGetErrorString(e));
        }
        cudaMemsetAsync(d_cc.p, 0, sizeof(ComplexCounters), s);
        int rc = RIN_OK;
        if (n) {
            switch (words_for(c->run_F)) {
            case 1: rc = get_complexes_w<1>(c, mode, d_req.as<uint32_t>(), (uint32_t)n, c->cx_out.as<uint32_t>(), (uint32_t)std::min<size_t>(cap_words, 0xffffffffu), d_span.as<uint2>(), d_cc.as<ComplexCounters>()); break;
            case 2: rc = get_complexes_w<2>(c, mode, d_req.as<uint32_t>(), (uint32_t)n, c->cx_out.as<uint32_t>(), (uint32_t)std::min<size_t>(cap_words, 0xffffffffu), d_span.as<uint2>(), d_cc.as<ComplexCounters>()); break;
            case 3: rc = get_complexes_w<3>(c, mode, d_req.as<uint32_t>(), (uint32_t)n, c->cx_out.as<uint32_t>(), (uint32_t)std::min<size_t>(cap_words, 0xffffffffu), d_span.as<uint2>(), d_cc.as<ComplexCounters>()); break;
            default: rc = get_complexes_w<4>(c, mode, d_req.as<uint32_t>(), (uint32_t)n, c->cx_out.as<uint32_t>(), (uint32_t)std::min<size_t>(cap_words, 0xffffffffu), d_span.as<uint2>(), d_cc.as<ComplexCounters>()); break;
            }
        }
        if (rc) {
            cleanup();
            return rc;
        }
        cudaMemcpyAsync(&hc, d_cc.p, sizeof(hc), cudaMemcpyDeviceToHost, s);
        if (n) cudaMemcpyAsync(span.data(), d_span.p, n * 8, cudaMemcpyDeviceToHost, s);
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) {
            cleanup();
            return fail(RIN_ERR_CUDA, cudaGetErrorString(e));
        }
        if (hc.err) {
            cleanup();
            return fail(hc.err, "rin_get_complexes: per-tet computation failed in tet " + std::to_string(hc.err_tet));
        }
        if (!hc.overflow) break;
        if (attempt > 2) {
            cleanup();
            return fail(RIN_ERR_STATE, "rin_get_complexes: output overflow");
        }
        cap_words = (size_t)hc.top + hc.top / 8 + 1024;
    }
    // repack in request order
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) total += span[i].y;
    *n_words = total;
    if (offsets) {
        uint64_t p = 0;
        for (uint64_t i = 0; i < n; ++i) {
            offsets[i] = p;
            p += span[i].y;
        }
        offsets[n] = p;
    }
    if (words && total) {
        std::vector<uint32_t> all(hc.top);
        cudaMemcpyAsync(all.data(), c->cx_out.p, (size_t)hc.top * 4, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        uint64_t p = 0;
        for (uint64_t i = 0; i < n; ++i) {
            memcpy(words + p, all.data() + span[i].x, (size_t)span[i].y * 4);
            p += span[i].y;
        }
    }
    cleanup();
    return RIN_OK;
}

extern "C++" {
namespace {
template <int W>
int robust_w(rin_ctx* c, int mode, RobustCounters* d_rc)
{
    const uint32_t A = (uint32_t)c->counts.num_intersecting_tet;
    if (!A) return RIN_OK;
    if (mode == RIN_MODE_IA)
        robust_test_kernel<W, false><<<grid_for(A, GEN_THREADS, c->sm_count, 4), GEN_THREADS, 0, c->stream>>>(
            c->tets.as<uint4>(), c->act_tet.as<uint32_t>(), c->act_mask.as<uint32_t>(), c->act_cap, A,
            c->vals.as<double>(), c->run_VS, d_rc);
    else
        robust_test_kernel<W, true><<<grid_for(A, GEN_THREADS, c->sm_count, 4), GEN_THREADS, 0, c->stream>>>(
            c->tets.as<uint4>(), c->act_tet.as<uint32_t>(), c->act_mask.as<uint32_t>(), c->act_cap, A,
            c->vals.as<double>(), c->run_VS, d_rc);
    CK(cudaGetLastError());
    return RIN_OK;
}
} // namespace
} // extern "C++"

// robust_test of the last run's active tets; out = {type1, type2, type3, tested}
int rin_tet_maps(rin_ctx* c, uint64_t* n_active, uint64_t* n_vert_entries, uint64_t* n_face_entries)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (!c->ran) return fail(RIN_ERR_STATE, "rin_tet_maps: no finished run");
    const bool mi = c->last_mode == RIN_MODE_MI;
    const uint8_t* blob = mi ? c->lut_mi.blob.as<uint8_t>() : c->lut_ia.blob.as<uint8_t>();
    if (c->marked || c->finalized) return fail(RIN_ERR_STATE, "rin_tet_maps: not available after a sharded exchange");
    if (c->t_count != c->own_count) return fail(RIN_ERR_STATE, "rin_tet_maps: not available for a run with ghost tets");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t A = (uint32_t)c->counts.num_intersecting_tet;
    if (!c->maps_ready) {
        CK(c->m_cnt.ensure((size_t)std::max(A, 1u) * 8));
        CK(c->m_off.ensure((size_t)(A + 1) * 8));
        uint2 tot = make_uint2(0, 0);
        if (A) {
            if (mi)
                tetmap_mi_count_kernel<<<grid_for(A, 256, c->sm_count), 256, 0, s>>>(c->rec_ref.as<uint32_t>(), A,
                    blob, c->arena.as<uint8_t>(), c->m_cnt.as<uint2>());
            else
                tetmap_count_kernel<<<grid_for(A, 256, c->sm_count), 256, 0, s>>>(c->rec_ref.as<uint32_t>(), A,
                    blob, c->arena.as<uint8_t>(), c->m_cnt.as<uint2>());
            scan_pairs_kernel<<<1, 1024, 0, s>>>(c->m_cnt.as<uint2>(), A, c->m_off.as<uint2>());
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&tot, c->m_off.as<uint2>() + A, 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        } else
            CK(cudaMemsetAsync(c->m_off.p, 0, 8, s));
        c->m_nv = tot.x;
        c->m_nf = tot.y;
        CK(c->m_vmap.ensure(std::max<size_t>(tot.x, 1) * 8));
        CK(c->m_fmap.ensure(std::max<size_t>(tot.y, 1) * 4));
        if (A) {
            const uint32_t* frep = c->ia_bndry_faces ? c->frep.as<uint32_t>() : nullptr;
            const uint4* fpos = c->ia_bndry_faces ? c->fpos.as<uint4>() : nullptr;
            if (mi)
                tetmap_mi_write_kernel<<<grid_for(A, 256, c->sm_count), 256, 0, s>>>(c->tets.as<uint4>(),
                    c->act_tet.as<uint32_t>(), A, c->rec_ref.as<uint32_t>(), c->offs.as<uint4>(), blob,
                    c->arena.as<uint8_t>(), c->rep.as<uint32_t>(), c->vid.as<uint32_t>(), frep, fpos,
                    c->fpartner.as<uint32_t>(), c->m_off.as<uint2>(), c->m_vmap.as<long long>(),
                    c->m_fmap.as<uint32_t>());
            else
                tetmap_write_kernel<<<grid_for(A, 256, c->sm_count), 256, 0, s>>>(c->tets.as<uint4>(),
                    c->act_tet.as<uint32_t>(), A, c->rec_ref.as<uint32_t>(), c->offs.as<uint4>(), blob,
                    c->arena.as<uint8_t>(), c->slot_of.as<uint32_t>(), c->table.as<uint32_t>(), frep, fpos,
                    c->m_off.as<uint2>(), c->m_vmap.as<long long>(), c->m_fmap.as<uint32_t>());
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(s));
        }
        c->maps_ready = true;
    }
    if (n_active) *n_active = A;
    if (n_vert_entries) *n_vert_entries = c->m_nv;
    if (n_face_entries) *n_face_entries = c->m_nf;
    return RIN_OK;
}

int rin_download_tet_maps(rin_ctx* c, uint32_t* active_tets, uint32_t* vert_offsets, int64_t* vert_ids,
    uint32_t* face_offsets, uint32_t* face_ids)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (!c->maps_ready) return fail(RIN_ERR_STATE, "rin_download_tet_maps: call rin_tet_maps first");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t A = (uint32_t)c->counts.num_intersecting_tet;
    if (active_tets && A) CK(cudaMemcpyAsync(active_tets, c->act_tet.p, (size_t)A * 4, cudaMemcpyDeviceToHost, s));
    if (vert_offsets || face_offsets) {
        std::vector<uint2> off(A + 1);
        CK(cudaMemcpyAsync(off.data(), c->m_off.p, (size_t)(A + 1) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (uint32_t a = 0; a <= A; ++a) {
            if (vert_offsets) vert_offsets[a] = off[a].x;
            if (face_offsets) face_offsets[a] = off[a].y;
        }
    }
    if (vert_ids && c->m_nv) CK(cudaMemcpyAsync(vert_ids, c->m_vmap.p, c->m_nv * 8, cudaMemcpyDeviceToHost, s));
    if (face_ids && c->m_nf) CK(cudaMemcpyAsync(face_ids, c->m_fmap.p, c->m_nf * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return RIN_OK;
}

// N1: compute_mesh_edges (src/mesh_connectivity.cpp:10-56) of the last run's mesh on the device (edges.cuh)
int rin_mesh_edges(rin_ctx* c, uint64_t* n_edges)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (!c->ran) return fail(RIN_ERR_STATE, "rin_mesh_edges: no finished run");
    CK(cudaSetDevice(c->device));
    if (!c->edges_ready) {
        cudaStream_t s = c->stream;
        const int sm = c->sm_count;
        const uint32_t NF = (uint32_t)c->counts.num_faces, NP = (uint32_t)c->counts.num_face_verts;
        c->n_edges = 0;
        if (NP) {
            uint32_t tsize = 1024;
            while (tsize < 2ull * NP) tsize <<= 1;
            CK(c->e_key.ensure((size_t)NP * 8));
            CK(c->e_slot.ensure((size_t)NP * 4));
            CK(c->e_table.ensure((size_t)tsize * 4));
            CK(c->e_verts.ensure((size_t)NP * 8));
            CK(c->e_of_face.ensure((size_t)NP * 4));
            CK(c->e_cnt.ensure((size_t)NP * 8 + 64)); // counts | cursors
            CK(c->e_off.ensure((size_t)(NP + 1) * 4));
            CK(c->e_pairs.ensure((size_t)NP * 8));
            const uint32_t tiles = (NP + 1023) / 1024;
            CK(c->status.ensure((size_t)tiles * 8 + 64));
            CK(c->counters.ensure(sizeof(Counters)));
            Counters* dctr = c->counters.as<Counters>();
            CK(cudaMemsetAsync(c->e_table.p, 0xff, (size_t)tsize * 4, s));
            CK(cudaMemsetAsync(c->e_cnt.p, 0, (size_t)NP * 8, s));
            CK(cudaMemsetAsync(c->status.p, 0, (size_t)tiles * 8, s));
            CK(cudaMemsetAsync(&dctr->rank_tile, 0, 8, s)); // rank_tile, n_unique
            edge_keys_kernel<<<grid_for(NF, 256, sm), 256, 0, s>>>(c->f_off.as<uint32_t>(), c->f_verts.as<uint32_t>(), NF,
                c->e_key.as<uint2>());
            edge_insert_kernel<<<grid_for(NP, 256, sm), 256, 0, s>>>(c->e_key.as<uint2>(), NP, c->e_table.as<uint32_t>(),
                tsize - 1, c->e_slot.as<uint32_t>());
            edge_rank_kernel<<<grid_for(tiles, 1, sm, 6), 256, 0, s>>>(c->e_key.as<uint2>(), NP,
                c->e_table.as<uint32_t>(), c->e_slot.as<uint32_t>(), c->e_verts.as<uint32_t>(),
                c->status.as<unsigned long long>(), &dctr->rank_tile, &dctr->n_unique);
            edge_assign_kernel<<<grid_for(NP, 256, sm), 256, 0, s>>>(NP, c->e_table.as<uint32_t>(),
                c->e_slot.as<uint32_t>(), c->e_of_face.as<uint32_t>(), c->e_cnt.as<uint32_t>());
            CK(cudaGetLastError());
            unsigned ne = 0;
            CK(cudaMemcpyAsync(&ne, &dctr->n_unique, 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            const uint32_t etiles = (ne + 1023) / 1024;
            CK(cudaMemsetAsync(c->status.p, 0, (size_t)std::max(etiles, 1u) * 8, s));
            CK(cudaMemsetAsync(&dctr->rank_tile, 0, 4, s));
            scan_u32_kernel<<<grid_for(std::max(etiles, 1u), 1, sm, 6), 256, 0, s>>>(c->e_cnt.as<uint32_t>(), ne,
                c->e_off.as<uint32_t>(), c->status.as<unsigned long long>(), &dctr->rank_tile);
            edge_fill_kernel<<<grid_for(NF, 256, sm), 256, 0, s>>>(c->f_off.as<uint32_t>(), NF,
                c->e_of_face.as<uint32_t>(), c->e_off.as<uint32_t>(), c->e_cnt.as<uint32_t>() + NP,
                c->e_pairs.as<uint2>());
            edge_sort_kernel<<<grid_for(ne, 256, sm), 256, 0, s>>>(ne, c->e_off.as<uint32_t>(), c->e_pairs.as<uint2>());
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(s));
            c->n_edges = ne;
        }
        c->edges_ready = true;
    }
    if (n_edges) *n_edges = c->n_edges;
    return RIN_OK;
}

int rin_download_edges(rin_ctx* c, rin_edges_out* o)
{
    if (!c || !o) return fail(RIN_ERR_ARG, "null argument");
    if (!c->edges_ready) return fail(RIN_ERR_STATE, "rin_download_edges: call rin_mesh_edges first");
    CK(cudaSetDevice(c->device));
    const size_t NE = c->n_edges, NP = (size_t)c->counts.num_face_verts;
    auto dl = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t {
        if (!dst || bytes == 0) return cudaSuccess;
        return cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    CK(dl(o->edge_verts, c->e_verts, NE * 8));
    CK(dl(o->edges_of_face, c->e_of_face, NP * 4));
    if (o->edge_face_offsets) {
        if (NE)
            CK(dl(o->edge_face_offsets, c->e_off, (NE + 1) * 4));
        else
            o->edge_face_offsets[0] = 0;
    }
    CK(dl(o->edge_faces, c->e_pairs, NP * 8));
    CK(cudaStreamSynchronize(c->stream));
    return RIN_OK;
}

int rin_robust_test(rin_ctx* c, int mode, uint32_t out[4])
{
    if (!c || !out) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran || c->last_mode != mode) return fail(RIN_ERR_STATE, "rin_robust_test: no finished run of that mode");
    CK(cudaSetDevice(c->device));
    DevBuf d;
    CK(d.ensure(sizeof(RobustCounters)));
    cudaMemsetAsync(d.p, 0, sizeof(RobustCounters), c->stream);
    int rc;
    switch (words_for(c->run_F)) {
    case 1: rc = robust_w<1>(c, mode, d.as<RobustCounters>()); break;
    case 2: rc = robust_w<2>(c, mode, d.as<RobustCounters>()); break;
    case 3: rc = robust_w<3>(c, mode, d.as<RobustCounters>()); break;
    default: rc = robust_w<4>(c, mode, d.as<RobustCounters>()); break;
    }
    if (rc == RIN_OK) {
        cudaMemcpyAsync(out, d.p, 16, cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(RIN_ERR_CUDA, "rin_robust_test: kernel failed");
    }
    d.release();
    return rc;
}

int rin_get_vertex_range(const rin_ctx* c, uint32_t* lo, uint32_t* hi)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (lo) *lo = c->v_count ? c->v_first : 0;
    if (hi) *hi = c->v_count ? c->v_first + c->v_count - 1 : (uint32_t)c->V - 1;
    return RIN_OK;
}

int rin_boundary_export(rin_ctx* c, int own_only, uint32_t lo, uint32_t hi, uint32_t* keys, uint32_t* ids,
    uint64_t capacity, uint64_t* n_out)
{
    if (!c || !n_out) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran) return fail(RIN_ERR_STATE, "no finished run");
    if (own_only && !c->marked) return fail(RIN_ERR_STATE, "rin_boundary_export(own_only): call rin_mark_foreign first");
    if (c->finalized) return fail(RIN_ERR_STATE, "run already finalized");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t NV = c->n_local_verts;
    CK(c->counters.ensure(sizeof(Counters)));
    unsigned* d_n = &c->counters.as<Counters>()->n_unique;
    CK(cudaMemsetAsync(d_n, 0, 4, s));
    const uint32_t cap = (uint32_t)std::max<uint64_t>(NV, 1);
    CK(c->bkeys.ensure((size_t)cap * 16));
    CK(c->bids.ensure((size_t)cap * 4));
    if (NV) {
        boundary_select_kernel<<<grid_for(NV, 256, c->sm_count), 256, 0, s>>>(c->v_key.as<uint4>(),
            c->v_size.as<uint8_t>(), NV, lo, hi, own_only ? c->own_idx.as<uint32_t>() : nullptr,
            c->bkeys.as<uint4>(), c->bids.as<uint32_t>(), cap, d_n);
        CK(cudaGetLastError());
    }
    unsigned n = 0;
    CK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *n_out = n;
    if (keys && ids) {
        if (capacity < n) return fail(RIN_ERR_ARG, "rin_boundary_export: capacity too small");
        if (n) {
            CK(cudaMemcpyAsync(keys, c->bkeys.p, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(ids, c->bids.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
    }
    return RIN_OK;
}

namespace {
int upload_foreign(rin_ctx* c, const uint32_t* keys, const uint32_t* gids, uint64_t m, uint32_t* mask_out)
{
    cudaStream_t s = c->stream;
    *mask_out = 0;
    if (m == 0) return RIN_OK;
    uint32_t tsize = 64;
    while (tsize < 2 * m) tsize <<= 1;
    CK(c->fkeys.ensure(m * 16));
    CK(c->ftable.ensure((size_t)tsize * 4));
    CK(cudaMemcpyAsync(c->fkeys.p, keys, m * 16, cudaMemcpyHostToDevice, s));
    if (gids) {
        CK(c->fgids.ensure(m * 4));
        CK(cudaMemcpyAsync(c->fgids.p, gids, m * 4, cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemsetAsync(c->ftable.p, 0xff, (size_t)tsize * 4, s));
    foreign_insert_kernel<<<grid_for(m, 256, c->sm_count), 256, 0, s>>>(c->fkeys.as<uint4>(), (uint32_t)m,
        c->ftable.as<uint32_t>(), tsize - 1);
    CK(cudaGetLastError());
    *mask_out = tsize - 1;
    return RIN_OK;
}
} // namespace

int rin_mark_foreign(rin_ctx* c, const uint32_t* keys, uint64_t m, uint64_t* n_own)
{
    if (!c || (m && !keys)) return fail(RIN_ERR_ARG, "null argument");
    if (!c->ran || c->finalized) return fail(RIN_ERR_STATE, "rin_mark_foreign: no fresh run");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t NV = c->n_local_verts;
    uint32_t mask = 0;
    int rc = upload_foreign(c, keys, nullptr, m, &mask);
    if (rc) return rc;
    CK(c->own_flag.ensure((size_t)std::max(NV, 1u) * 4));
    CK(c->own_idx.ensure((size_t)std::max(NV, 1u) * 4));
    unsigned* d_n = &c->counters.as<Counters>()->n_unique;
    if (NV) {
        mark_foreign_kernel<<<grid_for(NV, 256, c->sm_count), 256, 0, s>>>(c->v_key.as<uint4>(),
            c->v_size.as<uint8_t>(), NV, c->fkeys.as<uint4>(), c->ftable.as<uint32_t>(), mask,
            c->own_flag.as<uint32_t>(), c->ghost_v_lo);
        own_scan_kernel<<<1, 1024, 0, s>>>(c->own_flag.as<uint32_t>(), NV, c->own_idx.as<uint32_t>(), d_n);
        CK(cudaGetLastError());
    } else
        CK(cudaMemsetAsync(d_n, 0, 4, s));
    unsigned n = 0;
    CK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    c->n_own = n;
    c->marked = true;
    if (n_own) *n_own = n;
    return RIN_OK;
}

int rin_finalize_sharded(rin_ctx* c, uint64_t vert_offset, const uint32_t* keys, const uint32_t* gids, uint64_t m)
{
    if (!c || (m && (!keys || !gids))) return fail(RIN_ERR_ARG, "null argument");
    if (!c->marked || c->finalized) return fail(RIN_ERR_STATE, "rin_finalize_sharded: call rin_mark_foreign first");
    if (vert_offset + c->n_own >= 0xffffffffull) return fail(RIN_ERR_ARG, "global vertex ids exceed 32 bits");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t NV = c->n_local_verts, NO = c->n_own;
    uint32_t mask = 0;
    int rc = upload_foreign(c, keys, gids, m, &mask);
    if (rc) return rc;
    CK(c->gid.ensure((size_t)std::max(NV, 1u) * 4));
    unsigned* d_bad = &c->counters.as<Counters>()->n_bndry_faces;
    CK(cudaMemsetAsync(d_bad, 0, 4, s));
    if (NV) {
        global_ids_kernel<<<grid_for(NV, 256, c->sm_count), 256, 0, s>>>(c->v_key.as<uint4>(),
            c->own_idx.as<uint32_t>(), NV, (uint32_t)vert_offset, c->fkeys.as<uint4>(), c->fgids.as<uint32_t>(),
            c->ftable.as<uint32_t>(), mask, c->gid.as<uint32_t>(), d_bad, c->ghost_v_lo);
        const uint32_t NFV = (uint32_t)c->counts.num_face_verts;
        if (NFV && c->ghost_v_lo)
            check_gids_kernel<<<grid_for(NFV, 256, c->sm_count), 256, 0, s>>>(c->f_verts.as<uint32_t>(), NFV,
                c->gid.as<uint32_t>(), d_bad);
        if (NFV)
            apply_gids_kernel<<<grid_for(NFV, 256, c->sm_count), 256, 0, s>>>(c->f_verts.as<uint32_t>(), NFV,
                c->gid.as<uint32_t>());
        const size_t no1 = std::max(NO, 1u);
        CK(c->o_tet.ensure(no1 * 4));
        CK(c->o_local.ensure(no1));
        CK(c->o_size.ensure(no1));
        CK(c->o_simplex.ensure(no1 * 16));
        CK(c->o_funcs.ensure(no1 * 16));
        CK(c->o_xyz.ensure(no1 * 24));
        CK(c->o_key.ensure(no1 * 16));
        compact_own_verts_kernel<<<grid_for(NV, 256, c->sm_count), 256, 0, s>>>(c->own_idx.as<uint32_t>(), NV,
            c->v_tet.as<uint32_t>(), c->v_local.as<uint8_t>(), c->v_size.as<uint8_t>(), c->v_simplex.as<uint4>(),
            c->v_funcs.as<uint4>(), c->v_xyz.as<double>(), c->v_key.as<uint4>(), c->o_tet.as<uint32_t>(),
            c->o_local.as<uint8_t>(), c->o_size.as<uint8_t>(), c->o_simplex.as<uint4>(), c->o_funcs.as<uint4>(),
            c->o_xyz.as<double>(), c->o_key.as<uint4>());
        CK(cudaGetLastError());
    }
    unsigned bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (bad) return fail(RIN_ERR_STATE, "rin_finalize_sharded: " + std::to_string(bad) + " shared vertices have no owner id");
    std::swap(c->v_tet, c->o_tet);
    std::swap(c->v_local, c->o_local);
    std::swap(c->v_size, c->o_size);
    std::swap(c->v_simplex, c->o_simplex);
    std::swap(c->v_funcs, c->o_funcs);
    std::swap(c->v_xyz, c->o_xyz);
    std::swap(c->v_key, c->o_key);
    c->counts.num_verts = NO;
    c->finalized = true;
    return RIN_OK;
}

// ---- NCCL (loaded at run time; the process normally has torch's bundled libnccl.so.2 mapped) ----
namespace {


int load_nccl()
{
    if (g_nccl.handle) return RIN_OK;
    const char* env = getenv("RIN_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(RIN_ERR_STATE, "NCCL library not found (set RIN_NCCL_LIB)");
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    void* init = dlsym(h, "ncclCommInitRank");
    if (!g_nccl.GetUniqueId || !g_nccl.AllGather || !init || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd)
        return fail(RIN_ERR_STATE, "NCCL symbols missing");
    memcpy(&g_nccl.CommInitRank, &init, sizeof(init));
    g_nccl.handle = h;
    return RIN_OK;
}
#define NK(call)                                                                                   \
    do {                                                                                           \
        int r_ = (call);                                                                           \
        if (r_ != 0)                                                                               \
            return fail(RIN_ERR_CUDA, std::string(#call) + ": NCCL error " +                       \
                                          (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?")); \
    } while (0)
} // namespace

int rin_nccl_unique_id(uint8_t id[128])
{
    int rc = load_nccl();
    if (rc) return rc;
    NK(g_nccl.GetUniqueId(id));
    return RIN_OK;
}

int rin_nccl_init(rin_ctx* c, const uint8_t id[128], int rank, int world)
{
    if (!c || !id) return fail(RIN_ERR_ARG, "null argument");
    int rc = load_nccl();
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    typedef int (*init_fn)(void**, int, UniqueId, int);
    init_fn init;
    memcpy(&init, &g_nccl.CommInitRank, sizeof(init));
    UniqueId u;
    memcpy(u.internal, id, 128);
    if (c->nccl_comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy(c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    NK(init(&c->nccl_comm, world, u, rank));
    c->x_rank = rank;
    c->x_world = world;
    c->x_window = false;
    release_peers(c);
    c->p_failed = false;
    return RIN_OK;
}

// Neighbour protocol for slab sharding: rank r shares vertices with r-1 and r+1 only and owns everything on the
// plane it shares with r+1 (first occurrence in tet order).  Per pass: one ncclSend/ncclRecv pair with the keys of
// the upper plane and one ncclAllGather of (counts, own indices of those keys) for the global offsets and ids;
// all stream-ordered, ONE host synchronisation at the end.
extern "C++" {
namespace {
constexpr int X_NEED_GHOST = 1000; // internal: a rank saw degenerate vertices, the run has to be repeated with ghost tets
uint32_t need_ghost(const rin_ctx* c)
{
    return (c->counts.num_degenerate_vertex != 0 && c->ghost_lo == 0 && c->ghost_hi == 0) ? 1u : 0u;
}
int exchange_neighbours(rin_ctx* c, uint64_t* vert_offset, uint64_t* n_verts_total, uint64_t* face_offset,
    uint64_t* n_faces_total)
{
    cudaStream_t s = c->stream;
    const int rank = c->x_rank, world = c->x_world, sm = c->sm_count;
    const uint32_t NV = c->n_local_verts;
    const uint32_t NFV = (uint32_t)c->counts.num_face_verts;
    uint32_t* small = c->x_small.as<uint32_t>(); // [0..15] counters, [64..] offsets
    for (int attempt = 0;; ++attempt) {
        const uint32_t cap = c->x_cap;
        const size_t words = xmsg_words(cap);
        CK(c->x_send.ensure(words * 4));
        CK(c->x_recv1.ensure(words * 4));
        const size_t rec_words = 8 + (size_t)cap; // count record + own indices of the vertices sent upwards
        CK(c->x_cnt.ensure((size_t)(world + 1) * rec_words * 4));
        uint32_t tsize = 64;
        while (tsize < 2ull * cap) tsize <<= 1;
        CK(c->x_table.ensure((size_t)tsize * 4));
        CK(c->own_idx.ensure((size_t)std::max(NV, 1u) * 4));
        CK(c->gid.ensure((size_t)std::max(NV, 1u) * 4));
        const uint32_t tiles = (NV + 1023) / 1024;
        CK(c->status.ensure((size_t)std::max(tiles, 1u) * 8 + 64));
        const size_t no1 = std::max(NV, 1u); // owned vertices <= local vertices
        CK(c->o_tet.ensure(no1 * 4));
        CK(c->o_local.ensure(no1));
        CK(c->o_size.ensure(no1));
        CK(c->o_simplex.ensure(no1 * 16));
        CK(c->o_funcs.ensure(no1 * 16));
        CK(c->o_xyz.ensure(no1 * 24));
        CK(c->o_key.ensure(no1 * 16));
        CK(cudaMemsetAsync(small, 0, 64, s));
        unsigned* d_up = small + 0;
        unsigned* d_nown = small + 2;
        unsigned* d_ovf = small + 3;
        unsigned* d_tile = small + 4;
        unsigned* d_bad = small + 5;
        uint32_t* d_voff = small + 64;
        uint32_t* d_foff = small + 64 + (world + 1);
        uint32_t* up = c->x_send.as<uint32_t>();
        uint32_t* low = c->x_recv1.as<uint32_t>();
        uint32_t* mine4 = c->x_cnt.as<uint32_t>(); // this rank's record
        uint32_t* all4 = mine4 + rec_words;        // gathered
        // 1. keys (+ local indices) of the vertices on the plane shared with r+1 -> r+1
        if (NV && c->x_up_lo <= c->x_up_hi)
            x_select_kernel<<<grid_for(NV, 256, sm), 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), NV,
                c->x_up_lo, c->x_up_hi, nullptr, up, cap, d_up);
        x_header_kernel<<<1, 1, 0, s>>>(up, d_up, cap, nullptr, 0, d_ovf);
        NK(g_nccl.GroupStart());
        if (c->x_has_up) NK(g_nccl.Send(up, words, 3 /*ncclUint32*/, rank + 1, c->nccl_comm, s));
        if (c->x_has_low) NK(g_nccl.Recv(low, words, 3, rank - 1, c->nccl_comm, s));
        NK(g_nccl.GroupEnd());
        // 2. what the lower neighbour found first is foreign; ordered own index of the rest
        CK(cudaMemsetAsync(c->x_table.p, 0xff, (size_t)tsize * 4, s));
        if (c->x_has_low)
            x_insert_kernel<<<grid_for(cap, 256, sm), 256, 0, s>>>(low, cap, 1, c->x_table.as<uint32_t>(), tsize - 1);
        CK(cudaMemsetAsync(c->status.p, 0, (size_t)std::max(tiles, 1u) * 8, s));
        if (NV)
            x_mark_scan_kernel<<<tiles, 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), NV, low, cap,
                c->x_table.as<uint32_t>(), tsize - 1, c->x_has_low ? 1 : 0, c->own_idx.as<uint32_t>(),
                c->status.as<unsigned long long>(), d_tile, d_nown, c->ghost_v_lo);
        // 3. own indices of what was sent upwards -> r+1; (n_own, n_faces, n_up) of every rank -> offsets
        x_own_ids_kernel<<<grid_for(cap, 256, sm), 256, 0, s>>>(up, cap, c->own_idx.as<uint32_t>(), mine4, d_bad,
            d_nown, (uint32_t)c->counts.num_faces, (uint32_t)c->counts.num_face_verts,
            (uint32_t)c->counts.num_face_tets, need_ghost(c));
        NK(g_nccl.AllGather(mine4, all4, rec_words, 3, c->nccl_comm, s));
        x_offsets_nb_kernel<<<1, 1, 0, s>>>(all4, rec_words, world, cap, d_voff, d_foff, d_ovf, small + 6);
        const uint32_t* ids_low = all4 + (size_t)std::max(rank - 1, 0) * rec_words + 8;
        // 4. global ids, face vertex lists, owned vertices (skipped by every rank alike when a message overflowed:
        //    the decision comes from the gathered counts)
        if (NV) {
            x_global_ids_nb_kernel<<<grid_for(NV, 256, sm), 256, 0, s>>>(c->v_key.as<uint4>(),
                c->own_idx.as<uint32_t>(), NV, rank, d_voff, low, ids_low, cap,
                c->x_table.as<uint32_t>(), tsize - 1, c->gid.as<uint32_t>(), d_bad, c->ghost_v_lo);
            if (NFV && c->ghost_v_lo)
                check_gids_kernel<<<grid_for(NFV, 256, sm), 256, 0, s>>>(c->f_verts.as<uint32_t>(), NFV,
                    c->gid.as<uint32_t>(), d_bad);
            compact_own_verts_kernel<<<grid_for(NV, 256, sm), 256, 0, s>>>(c->own_idx.as<uint32_t>(), NV,
                c->v_tet.as<uint32_t>(), c->v_local.as<uint8_t>(), c->v_size.as<uint8_t>(), c->v_simplex.as<uint4>(),
                c->v_funcs.as<uint4>(), c->v_xyz.as<double>(), c->v_key.as<uint4>(), c->o_tet.as<uint32_t>(),
                c->o_local.as<uint8_t>(), c->o_size.as<uint8_t>(), c->o_simplex.as<uint4>(), c->o_funcs.as<uint4>(),
                c->o_xyz.as<double>(), c->o_key.as<uint4>());
        }
        std::vector<uint32_t> hsmall(64 + 4 * (world + 1));
        CK(cudaMemcpyAsync(hsmall.data(), small, hsmall.size() * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s)); // the one synchronisation of the exchange
        if (hsmall[3]) { // a rank's message outgrew the capacity: every rank sees it in the gathered counts
            if (attempt > 3) return fail(RIN_ERR_STATE, "exchange: capacity negotiation failed");
            c->x_cap = (hsmall[3] + hsmall[3] / 4 + 1024 + 3u) & ~3u;
            continue;
        }
        if (hsmall[6]) return X_NEED_GHOST; // every rank sees the same gathered flags
        if (hsmall[5])
            return fail(RIN_ERR_STATE, "exchange: " + std::to_string(hsmall[5]) + " shared vertices have no owner");
        // only now (nothing can fail any more) the face vertex lists are rewritten in place
        if (NV && NFV) {
            apply_gids_kernel<<<grid_for(NFV, 256, sm), 256, 0, s>>>(c->f_verts.as<uint32_t>(), NFV,
                c->gid.as<uint32_t>());
            CK(cudaGetLastError());
        }
        const uint32_t NO = hsmall[2];
        std::swap(c->v_tet, c->o_tet);
        std::swap(c->v_local, c->o_local);
        std::swap(c->v_size, c->o_size);
        std::swap(c->v_simplex, c->o_simplex);
        std::swap(c->v_funcs, c->o_funcs);
        std::swap(c->v_xyz, c->o_xyz);
        std::swap(c->v_key, c->o_key);
        c->n_own = NO;
        c->counts.num_verts = NO;
        c->marked = c->finalized = true;
        for (int q = 0; q < 4; ++q) {
            c->x_offsets[2 * q] = hsmall[64 + q * (world + 1) + rank];
            c->x_offsets[2 * q + 1] = hsmall[64 + q * (world + 1) + world];
        }
        {
            // face offset arrays -> offsets into the merged face_verts / face_tets arrays
            const uint32_t nf1 = (uint32_t)c->counts.num_faces + 1;
            add_offset_kernel<<<grid_for(nf1, 256, sm), 256, 0, s>>>(c->f_off.as<uint32_t>(), nf1,
                (uint32_t)c->x_offsets[4]);
            add_offset_kernel<<<grid_for(nf1, 256, sm), 256, 0, s>>>(c->f_toff.as<uint32_t>(), nf1,
                (uint32_t)c->x_offsets[6]);
            CK(cudaGetLastError());
        }
        if (vert_offset) *vert_offset = c->x_offsets[0];
        if (n_verts_total) *n_verts_total = c->x_offsets[1];
        if (face_offset) *face_offset = c->x_offsets[2];
        if (n_faces_total) *n_faces_total = c->x_offsets[3];
        return RIN_OK;
    }
}
} // namespace
} // extern "C++"

extern "C++" {
namespace {
int exchange_once(rin_ctx* c, uint64_t* vert_offset, uint64_t* n_verts_total, uint64_t* face_offset,
    uint64_t* n_faces_total);
}
}

// The whole slab-boundary protocol on the device (see sharding.py).  When any rank reports degenerate vertices
// and the run had no ghost tets, every rank (they all see the same gathered flags) widens its range by one cube
// layer of the generated grid, repeats the run and the exchange: iso-faces lying on the slab plane then get both
// incident tets and the material matching across the plane happens (rin_set_ghost_tets).
int rin_exchange_nccl(rin_ctx* c, uint64_t* vert_offset, uint64_t* n_verts_total, uint64_t* face_offset,
    uint64_t* n_faces_total)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    int rc = exchange_once(c, vert_offset, n_verts_total, face_offset, n_faces_total);
    if (rc != X_NEED_GHOST) return rc;
    if (c->grid_R == 0)
        return fail(RIN_ERR_STATE, "degenerate vertices in a sharded run on an unstructured mesh: give every rank the "
                                   "ghost tets that touch its shared vertices (rin_set_ghost_tets) and run again");
    const uint64_t layer = 5ull * c->grid_R * c->grid_R;
    const int mode = c->last_mode;
    const uint32_t flags = c->last_flags;
    rc = rin_set_ghost_tets(c, layer, layer);
    if (rc) return rc;
    rc = rin_run(c, mode, flags);
    if (rc) return rc;
    rc = exchange_once(c, vert_offset, n_verts_total, face_offset, n_faces_total);
    if (rc == X_NEED_GHOST) return fail(RIN_ERR_STATE, "exchange: ghost negotiation failed");
    return rc;
}

extern "C++" {
namespace {
int exchange_once(rin_ctx* c, uint64_t* vert_offset, uint64_t* n_verts_total, uint64_t* face_offset,
    uint64_t* n_faces_total)
{
    if (!c->nccl_comm) return fail(RIN_ERR_STATE, "rin_exchange_nccl: call rin_nccl_init first");
    if (!c->ran || c->finalized) return fail(RIN_ERR_STATE, "rin_exchange_nccl: no fresh run");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int rank = c->x_rank, world = c->x_world, sm = c->sm_count;
    const uint32_t NV = c->n_local_verts;
    const uint32_t NFV = (uint32_t)c->counts.num_face_verts;
    CK(c->x_small.ensure(4096 + 64 * (size_t)c->x_world));
    uint32_t* small = c->x_small.as<uint32_t>(); // [0..15] counters, [16..16+2*world] ranges, [64..] offsets
    if (!c->x_window) {
        // shared vertex window: all-gather of the ranks' vertex ranges (once per mesh / range)
        uint32_t lo = c->v_count ? c->v_first : 0, hi = c->v_count ? c->v_first + c->v_count - 1 : (uint32_t)c->V - 1;
        if (c->t_count == 0) { // an empty tet range touches no vertex
            lo = 1;
            hi = 0;
        }
        uint32_t mine[2] = {lo, hi};
        std::vector<uint32_t> all(2 * world);
        CK(cudaMemcpyAsync(small + 16, mine, 8, cudaMemcpyHostToDevice, s));
        NK(g_nccl.AllGather(small + 16, small + 32, 2, 3 /*ncclUint32*/, c->nccl_comm, s));
        CK(cudaMemcpyAsync(all.data(), small + 32, 8 * world, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        bool any = false;
        uint32_t wl = 0, wh = 0;
        for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            uint32_t a = std::max(lo, all[2 * r]), b = std::min(hi, all[2 * r + 1]);
            if (a > b) continue;
            wl = any ? std::min(wl, a) : a;
            wh = any ? std::max(wh, b) : b;
            any = true;
        }
        c->x_lo = any ? wl : 1;
        c->x_hi = any ? wh : 0;
        c->x_window = true;
        // the same decision on every rank (all see all ranges): do non-adjacent ranks ever share a vertex?
        bool nb = true;
        for (int a = 0; a < world && nb; ++a)
            for (int b = a + 2; b < world && nb; ++b)
                if (std::max(all[2 * a], all[2 * b]) <= std::min(all[2 * a + 1], all[2 * b + 1]) &&
                    all[2 * a] <= all[2 * a + 1] && all[2 * b] <= all[2 * b + 1])
                    nb = false;
        for (int a = 0; a + 1 < world && nb; ++a) // ranges must ascend with the rank (the lower rank owns)
            if (all[2 * a] <= all[2 * a + 1] && all[2 * a + 2] <= all[2 * a + 3] && all[2 * a] > all[2 * a + 2]) nb = false;
        c->x_neighbours = nb && getenv("RIN_X_ALLGATHER") == nullptr;
        bool none_empty = true;
        for (int a = 0; a < world; ++a) none_empty &= all[2 * a] <= all[2 * a + 1];
        c->x_fusable = c->x_neighbours && none_empty && getenv("RIN_NO_FUSED_EXCHANGE") == nullptr;
        c->fuse_disabled = false;
        c->x_has_low = c->x_has_up = false;
        c->x_up_lo = 1;
        c->x_up_hi = 0;
        if (rank + 1 < world) {
            const uint32_t a = std::max(lo, all[2 * (rank + 1)]), b = std::min(hi, all[2 * (rank + 1) + 1]);
            c->x_has_up = true; // the message is exchanged even when the window is empty (sizes must match)
            if (a <= b && lo <= hi) {
                c->x_up_lo = a;
                c->x_up_hi = b;
            }
        }
        if (rank > 0) c->x_has_low = true;
    }
    if (c->x_cap == 0) {
        const char* env = getenv("RIN_XCAP"); // test hook: a tiny capacity forces the renegotiation path
        c->x_cap = env ? (uint32_t)std::max(4, atoi(env)) : 4096;
    }
    c->x_cap = (c->x_cap + 3u) & ~3u; // keys are read as uint4: keep every rank's segment 16-byte aligned
    if (c->x_neighbours) return exchange_neighbours(c, vert_offset, n_verts_total, face_offset, n_faces_total);
    for (int attempt = 0;; ++attempt) {
        const uint32_t cap = c->x_cap;
        const size_t words = xmsg_words(cap);
        CK(c->x_send.ensure(words * 4));
        CK(c->x_recv1.ensure(words * 4 * world));
        CK(c->x_recv2.ensure(words * 4 * world));
        uint32_t tsize = 64;
        while (tsize < 2ull * cap * std::max(rank, 1)) tsize <<= 1;
        CK(c->x_table.ensure((size_t)tsize * 4));
        CK(c->own_idx.ensure((size_t)std::max(NV, 1u) * 4));
        CK(c->gid.ensure((size_t)std::max(NV, 1u) * 4));
        const uint32_t tiles = (NV + 1023) / 1024;
        CK(c->status.ensure((size_t)std::max(tiles, 1u) * 8 + 64));
        CK(cudaMemsetAsync(small, 0, 64, s));
        unsigned* d_cnt1 = small + 0;
        unsigned* d_cnt2 = small + 1;
        unsigned* d_nown = small + 2;
        unsigned* d_ovf = small + 3;
        unsigned* d_tile = small + 4;
        unsigned* d_bad = small + 5;
        uint32_t* d_voff = small + 64;
        uint32_t* d_foff = small + 64 + (world + 1);
        uint32_t* msg = c->x_send.as<uint32_t>();
        // 1. candidates in the shared window -> all ranks
        if (NV)
            x_select_kernel<<<grid_for(NV, 256, sm), 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), NV,
                c->x_lo, c->x_hi, nullptr, msg, cap, d_cnt1);
        x_header_kernel<<<1, 1, 0, s>>>(msg, d_cnt1, cap, nullptr, 0, d_ovf);
        NK(g_nccl.AllGather(msg, c->x_recv1.p, words, 3, c->nccl_comm, s));
        // 2. keys of lower ranks are foreign; ordered own index
        CK(cudaMemsetAsync(c->x_table.p, 0xff, (size_t)tsize * 4, s));
        if (rank > 0)
            x_insert_kernel<<<grid_for((uint64_t)rank * cap, 256, sm), 256, 0, s>>>(c->x_recv1.as<uint32_t>(), cap,
                rank, c->x_table.as<uint32_t>(), tsize - 1);
        CK(cudaMemsetAsync(c->status.p, 0, (size_t)std::max(tiles, 1u) * 8, s));
        if (NV)
            x_mark_scan_kernel<<<tiles, 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), NV,
                c->x_recv1.as<uint32_t>(), cap, c->x_table.as<uint32_t>(), tsize - 1, rank, c->own_idx.as<uint32_t>(),
                c->status.as<unsigned long long>(), d_tile, d_nown, c->ghost_v_lo);
        // 3. owned shared keys with their own index + (n_own, n_faces) -> all ranks
        if (NV)
            x_select_kernel<<<grid_for(NV, 256, sm), 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), NV,
                c->x_lo, c->x_hi, c->own_idx.as<uint32_t>(), msg, cap, d_cnt2);
        x_header_kernel<<<1, 1, 0, s>>>(msg, d_cnt2, cap, d_nown, (uint32_t)c->counts.num_faces, d_ovf,
            (uint32_t)c->counts.num_face_verts, (uint32_t)c->counts.num_face_tets, need_ghost(c));
        NK(g_nccl.AllGather(msg, c->x_recv2.p, words, 3, c->nccl_comm, s));
        x_offsets_kernel<<<1, 1, 0, s>>>(c->x_recv1.as<uint32_t>(), c->x_recv2.as<uint32_t>(), cap, world, d_voff, d_foff,
            d_ovf, small + 6);
        // 4. global ids, rewrite the face vertex lists, keep the owned vertices
        CK(cudaMemsetAsync(c->x_table.p, 0xff, (size_t)tsize * 4, s));
        if (rank > 0)
            x_insert_kernel<<<grid_for((uint64_t)rank * cap, 256, sm), 256, 0, s>>>(c->x_recv2.as<uint32_t>(), cap,
                rank, c->x_table.as<uint32_t>(), tsize - 1);
        std::vector<uint32_t> hsmall(64 + 4 * (world + 1));
        CK(cudaMemcpyAsync(hsmall.data(), small, hsmall.size() * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (hsmall[3]) { // a rank's payload outgrew the capacity: every rank sees it, redo larger
            if (attempt > 3) return fail(RIN_ERR_STATE, "exchange: capacity negotiation failed");
            c->x_cap = (hsmall[3] + hsmall[3] / 4 + 1024 + 3u) & ~3u;
            continue;
        }
        if (hsmall[6]) return X_NEED_GHOST;
        const uint32_t NO = hsmall[2];
        if (NV) {
            x_global_ids_kernel<<<grid_for(NV, 256, sm), 256, 0, s>>>(c->v_key.as<uint4>(), c->own_idx.as<uint32_t>(),
                NV, rank, d_voff, c->x_recv2.as<uint32_t>(), cap, c->x_table.as<uint32_t>(), tsize - 1,
                c->gid.as<uint32_t>(), d_bad, c->ghost_v_lo);
            if (NFV && c->ghost_v_lo)
                check_gids_kernel<<<grid_for(NFV, 256, sm), 256, 0, s>>>(c->f_verts.as<uint32_t>(), NFV,
                    c->gid.as<uint32_t>(), d_bad);
            if (NFV)
                apply_gids_kernel<<<grid_for(NFV, 256, sm), 256, 0, s>>>(c->f_verts.as<uint32_t>(), NFV,
                    c->gid.as<uint32_t>());
            const size_t no1 = std::max(NO, 1u);
            CK(c->o_tet.ensure(no1 * 4));
            CK(c->o_local.ensure(no1));
            CK(c->o_size.ensure(no1));
            CK(c->o_simplex.ensure(no1 * 16));
            CK(c->o_funcs.ensure(no1 * 16));
            CK(c->o_xyz.ensure(no1 * 24));
            CK(c->o_key.ensure(no1 * 16));
            compact_own_verts_kernel<<<grid_for(NV, 256, sm), 256, 0, s>>>(c->own_idx.as<uint32_t>(), NV,
                c->v_tet.as<uint32_t>(), c->v_local.as<uint8_t>(), c->v_size.as<uint8_t>(), c->v_simplex.as<uint4>(),
                c->v_funcs.as<uint4>(), c->v_xyz.as<double>(), c->v_key.as<uint4>(), c->o_tet.as<uint32_t>(),
                c->o_local.as<uint8_t>(), c->o_size.as<uint8_t>(), c->o_simplex.as<uint4>(), c->o_funcs.as<uint4>(),
                c->o_xyz.as<double>(), c->o_key.as<uint4>());
            CK(cudaGetLastError());
        }
        unsigned bad = 0;
        CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (bad) return fail(RIN_ERR_STATE, "exchange: " + std::to_string(bad) + " shared vertices have no owner");
        std::swap(c->v_tet, c->o_tet);
        std::swap(c->v_local, c->o_local);
        std::swap(c->v_size, c->o_size);
        std::swap(c->v_simplex, c->o_simplex);
        std::swap(c->v_funcs, c->o_funcs);
        std::swap(c->v_xyz, c->o_xyz);
        std::swap(c->v_key, c->o_key);
        c->n_own = NO;
        c->counts.num_verts = NO;
        c->marked = c->finalized = true;
        for (int q = 0; q < 4; ++q) {
            c->x_offsets[2 * q] = hsmall[64 + q * (world + 1) + rank];
            c->x_offsets[2 * q + 1] = hsmall[64 + q * (world + 1) + world];
        }
        {
            // face offset arrays -> offsets into the merged face_verts / face_tets arrays
            const uint32_t nf1 = (uint32_t)c->counts.num_faces + 1;
            add_offset_kernel<<<grid_for(nf1, 256, sm), 256, 0, s>>>(c->f_off.as<uint32_t>(), nf1,
                (uint32_t)c->x_offsets[4]);
            add_offset_kernel<<<grid_for(nf1, 256, sm), 256, 0, s>>>(c->f_toff.as<uint32_t>(), nf1,
                (uint32_t)c->x_offsets[6]);
            CK(cudaGetLastError());
        }
        if (vert_offset) *vert_offset = c->x_offsets[0];
        if (n_verts_total) *n_verts_total = c->x_offsets[1];
        if (face_offset) *face_offset = c->x_offsets[2];
        if (n_faces_total) *n_faces_total = c->x_offsets[3];
        return RIN_OK;
    }
}

} // namespace
} // extern "C++"


extern "C++" {
namespace {
constexpr int X_NOT_FUSED = 1001; // internal: degenerate inputs, every rank falls back to rin_run + rin_exchange_nccl

// second stream of the fused run + exchange, HIGH priority: its (small, latency-bound) kernels get the SM slots the
// face kernel's blocks free first, so the exchange really runs beside the face kernel instead of behind it
int ensure_stream2(rin_ctx* c)
{
    if (c->stream2) return RIN_OK;
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, greatest));
    CK(cudaEventCreate(&c->ev_fork));
    CK(cudaEventCreate(&c->ev_join));
    return RIN_OK;
}
#define CK_RC(call)            \
    do {                       \
        int rc__ = (call);     \
        if (rc__) return rc__; \
    } while (0)

// The neighbour exchange enqueued behind the kernels of a run whose counts are still on the device
// (exchange_neighbours with every count read from device memory).  vcap = capacity of the vertex arrays.
int enqueue_fused_exchange(rin_ctx* c, Counters* dctr, bool known_degenerate, uint32_t vcap)
{
    // forked behind the vertex kernel: everything up to the global ids needs the vertices and the totals only
    CK_RC(ensure_stream2(c));
    cudaStream_t s = c->stream2;
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(s, c->ev_fork, 0));
    const int rank = c->x_rank, world = c->x_world, sm = c->sm_count;
    const uint32_t cap = c->x_cap;
    const size_t words = xmsg_words(cap);
    CK(c->x_small.ensure(4096 + 64 * (size_t)world));
    CK(c->x_send.ensure(words * 4));
    CK(c->x_recv1.ensure(words * 4));
    const size_t rec_words = 8 + (size_t)cap;
    CK(c->x_cnt.ensure((size_t)(world + 1) * rec_words * 4));
    uint32_t tsize = 64;
    while (tsize < 2ull * cap) tsize <<= 1;
    CK(c->x_table.ensure((size_t)tsize * 4));
    const size_t nv1 = std::max(vcap, 1u);
    CK(c->own_idx.ensure(nv1 * 4));
    CK(c->gid.ensure(nv1 * 4));
    const uint32_t tiles = (vcap + 1023) / 1024;
    CK(c->x_status.ensure((size_t)std::max(tiles, 1u) * 8 + 64));
    CK(c->o_tet.ensure(nv1 * 4));
    CK(c->o_local.ensure(nv1));
    CK(c->o_size.ensure(nv1));
    CK(c->o_simplex.ensure(nv1 * 16));
    CK(c->o_funcs.ensure(nv1 * 16));
    CK(c->o_xyz.ensure(nv1 * 24));
    CK(c->o_key.ensure(nv1 * 16));
    if (!c->h_xsmall) CK(cudaHostAlloc((void**)&c->h_xsmall, 4096 + 64 * 64, cudaHostAllocDefault));
    uint32_t* small = c->x_small.as<uint32_t>();
    CK(cudaMemsetAsync(small, 0, 64, s));
    unsigned* d_up = small + 0;
    unsigned* d_nown = small + 2;
    unsigned* d_ovf = small + 3;
    unsigned* d_tile = small + 4;
    unsigned* d_bad = small + 5;
    uint32_t* d_voff = small + 64;
    uint32_t* d_foff = small + 64 + (world + 1);
    uint32_t* up = c->x_send.as<uint32_t>();
    uint32_t* low = c->x_recv1.as<uint32_t>();
    uint32_t* mine4 = c->x_cnt.as<uint32_t>();
    uint32_t* all4 = mine4 + rec_words;
    const unsigned* d_nv = &dctr->n_unique;
    const int gv = grid_for(vcap, 256, sm);
    if (c->x_up_lo <= c->x_up_hi)
        x_select_kernel<<<gv, 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), vcap, c->x_up_lo, c->x_up_hi,
            nullptr, up, cap, d_up, d_nv);
    x_header_kernel<<<1, 1, 0, s>>>(up, d_up, cap, nullptr, 0, d_ovf);
    NK(g_nccl.GroupStart());
    if (c->x_has_up) NK(g_nccl.Send(up, words, 3 /*ncclUint32*/, rank + 1, c->nccl_comm, s));
    if (c->x_has_low) NK(g_nccl.Recv(low, words, 3, rank - 1, c->nccl_comm, s));
    NK(g_nccl.GroupEnd());
    CK(cudaMemsetAsync(c->x_table.p, 0xff, (size_t)tsize * 4, s));
    if (c->x_has_low)
        x_insert_kernel<<<grid_for(cap, 256, sm), 256, 0, s>>>(low, cap, 1, c->x_table.as<uint32_t>(), tsize - 1);
    CK(cudaMemsetAsync(c->x_status.p, 0, (size_t)std::max(tiles, 1u) * 8, s));
    x_mark_scan_kernel<<<std::max(tiles, 1u), 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), vcap, low, cap,
        c->x_table.as<uint32_t>(), tsize - 1, c->x_has_low ? 1 : 0, c->own_idx.as<uint32_t>(),
        c->x_status.as<unsigned long long>(), d_tile, d_nown, 0u, d_nv);
    x_own_ids_kernel<<<grid_for(cap, 256, sm), 256, 0, s>>>(up, cap, c->own_idx.as<uint32_t>(), mine4, d_bad, d_nown, 0, 0,
        0, 0);
    fx_header_kernel<<<1, 1, 0, s>>>(mine4, &dctr->tot.n_faces, &dctr->tot.n_fv, &dctr->n_zero, known_degenerate ? 1u : 0u,
        &dctr->overflow, reinterpret_cast<const unsigned*>(&dctr->gen.err), &dctr->gen.arena_overflow,
        &dctr->gen.n_bnd_faces);
    NK(g_nccl.AllGather(mine4, all4, rec_words, 3, c->nccl_comm, s));
    x_offsets_nb_kernel<<<1, 1, 0, s>>>(all4, rec_words, world, cap, d_voff, d_foff, d_ovf, small + 6);
    const uint32_t* ids_low = all4 + (size_t)std::max(rank - 1, 0) * rec_words + 8;
    x_global_ids_nb_kernel<<<gv, 256, 0, s>>>(c->v_key.as<uint4>(), c->own_idx.as<uint32_t>(), vcap, rank, d_voff, low,
        ids_low, cap, c->x_table.as<uint32_t>(), tsize - 1, c->gid.as<uint32_t>(), d_bad, 0u, d_nv);
    compact_own_verts_kernel<<<gv, 256, 0, s>>>(c->own_idx.as<uint32_t>(), vcap, c->v_tet.as<uint32_t>(),
        c->v_local.as<uint8_t>(), c->v_size.as<uint8_t>(), c->v_simplex.as<uint4>(), c->v_funcs.as<uint4>(),
        c->v_xyz.as<double>(), c->v_key.as<uint4>(), c->o_tet.as<uint32_t>(), c->o_local.as<uint8_t>(),
        c->o_size.as<uint8_t>(), c->o_simplex.as<uint4>(), c->o_funcs.as<uint4>(), c->o_xyz.as<double>(),
        c->o_key.as<uint4>(), d_nv);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev_join, s));
    c->launches += 9;
    return RIN_OK;
}


// Collective: every rank allocates its inbox, the CUDA IPC handles travel through one ncclAllGather, every rank maps
// its peers' inboxes; a second all-gather tells every rank whether ALL mappings succeeded (otherwise every rank keeps
// the NCCL data path) and doubles as the barrier between "inbox zeroed" and "first remote store".
int peer_setup(rin_ctx* c)
{
    const int world = c->x_world, rank = c->x_rank;
    cudaStream_t s = c->stream;
    release_peers(c);
    c->p_failed = true; // until proven otherwise
    if (world > PX_MAX_WORLD || getenv("RIN_NO_PEER_EXCHANGE")) return RIN_OK;
    const uint32_t pcap = (std::max<uint32_t>(2 * c->x_cap, 16384) + 3u) & ~3u;
    const size_t bytes = px_inbox_words(pcap, world) * 4;
    bool ok = cudaMalloc((void**)&c->p_inbox, bytes) == cudaSuccess;
    cudaIpcMemHandle_t mine{};
    if (ok) ok = cudaMemsetAsync(c->p_inbox, 0, bytes, s) == cudaSuccess;
    if (ok) ok = cudaIpcGetMemHandle(&mine, c->p_inbox) == cudaSuccess;
    if (!ok) cudaGetLastError();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CK(c->p_stage.ensure((size_t)(world + 1) * 64 + (size_t)(world + 1) * 4));
    uint8_t* st = c->p_stage.as<uint8_t>();
    CK(cudaMemcpyAsync(st, &mine, 64, cudaMemcpyHostToDevice, s));
    NK(g_nccl.AllGather(st, st + 64, 16, 3 /*ncclUint32*/, c->nccl_comm, s));
    std::vector<cudaIpcMemHandle_t> all(world);
    CK(cudaMemcpyAsync(all.data(), st + 64, (size_t)world * 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int r = 0; r < world && ok; ++r) {
        if (r == rank) {
            c->p_peer[r] = c->p_inbox;
            continue;
        }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
        } else
            c->p_peer[r] = static_cast<uint32_t*>(ptr);
    }
    uint32_t flag = ok ? 1u : 0u;
    uint32_t* fl = reinterpret_cast<uint32_t*>(st + (size_t)(world + 1) * 64);
    CK(cudaMemcpyAsync(fl, &flag, 4, cudaMemcpyHostToDevice, s));
    NK(g_nccl.AllGather(fl, fl + 1, 1, 3, c->nccl_comm, s));
    std::vector<uint32_t> flags(world);
    CK(cudaMemcpyAsync(flags.data(), fl + 1, (size_t)world * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    bool all_ok = true;
    for (uint32_t f : flags) all_ok &= f != 0;
    if (!all_ok) {
        release_peers(c);
        return RIN_OK; // NCCL data path
    }
    c->p_cap = pcap;
    c->p_pass = 0;
    c->p_ready = true;
    c->p_failed = false;
    return RIN_OK;
}

// the exchange chain with peer-memory kernels instead of NCCL calls (see exchange.cuh); same products as
// enqueue_fused_exchange
int enqueue_peer_exchange(rin_ctx* c, Counters* dctr, bool known_degenerate, uint32_t vcap)
{
    CK_RC(ensure_stream2(c));
    cudaStream_t s = c->stream2;
    const int rank = c->x_rank, world = c->x_world, sm = c->sm_count;
    const uint32_t cap = c->x_cap;
    const uint32_t pass = ++c->p_pass, parity = pass & 1u;
    CK(c->x_small.ensure(4096 + 64 * (size_t)world));
    CK(c->x_send.ensure(std::max<size_t>(cap, 1) * 4)); // local ids of the vertices sent upwards
    uint32_t tsize = 64;
    while (tsize < 2ull * cap) tsize <<= 1;
    CK(c->x_table.ensure((size_t)tsize * 4));
    const size_t nv1 = std::max(vcap, 1u);
    CK(c->own_idx.ensure(nv1 * 4));
    CK(c->gid.ensure(nv1 * 4));
    const uint32_t tiles = (vcap + 1023) / 1024;
    CK(c->x_status.ensure((size_t)std::max(tiles, 1u) * 8 + 64));
    CK(c->o_tet.ensure(nv1 * 4));
    CK(c->o_local.ensure(nv1));
    CK(c->o_size.ensure(nv1));
    CK(c->o_simplex.ensure(nv1 * 16));
    CK(c->o_funcs.ensure(nv1 * 16));
    CK(c->o_xyz.ensure(nv1 * 24));
    CK(c->o_key.ensure(nv1 * 16));
    if (!c->h_xsmall) CK(cudaHostAlloc((void**)&c->h_xsmall, 4096 + 64 * 64, cudaHostAllocDefault));
    uint32_t* small = c->x_small.as<uint32_t>();
    // the clears run EARLY (the second stream is idle while the run's kernels execute), the chain itself starts
    // behind the vertex kernel
    CK(cudaMemsetAsync(small, 0, 64, s));
    CK(cudaMemsetAsync(c->x_table.p, 0xff, (size_t)tsize * 4, s));
    CK(cudaMemsetAsync(c->x_status.p, 0, (size_t)std::max(tiles, 1u) * 8, s));
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(s, c->ev_fork, 0));
    unsigned* d_up = small + 0;
    unsigned* d_nown = small + 2;
    unsigned* d_ovf = small + 3;
    unsigned* d_tile = small + 4;
    unsigned* d_bad = small + 5;
    unsigned* d_done1 = small + 8;
    unsigned* d_done2 = small + 9;
    unsigned* d_timeout = small + 10;
    uint32_t* d_voff = small + 64;
    uint32_t* d_foff = small + 64 + (world + 1);
    uint32_t* local_ids = c->x_send.as<uint32_t>();
    uint32_t* low = px_keys(c->p_inbox, cap, world, parity);
    const uint32_t* ids_low = px_record(c->p_inbox, cap, world, parity, std::max(rank - 1, 0)) + 8;
    PeerInboxes peers{};
    for (int r = 0; r < world; ++r) peers.p[r] = c->p_peer[r];
    const unsigned* d_nv = &dctr->n_unique;
    // grid-stride kernels over the (device-side) vertex count: a few blocks per SM are plenty, and every block of the
    // send / publish kernels pays a system-scope fence
    const int gv = grid_for(std::max<uint32_t>(c->h_unique, 4096), 256, sm, 2);
    int xe = 0;
    auto XEV = [&]() {
        if (c->stage_timing && xe < 10) {
            if (!c->ev_x[xe]) cudaEventCreate(&c->ev_x[xe]);
            cudaEventRecord(c->ev_x[xe++], s);
        }
    };
    px_send_keys_kernel<<<gv, 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), d_nv, c->x_up_lo, c->x_up_hi,
        c->x_has_up ? c->p_peer[rank + 1] : nullptr, cap, world, pass, local_ids, d_up, d_done1, d_ovf);
    XEV(); // 0: send keys
    if (c->x_has_low)
        px_recv_insert_kernel<<<grid_for(cap, 256, sm), 256, 0, s>>>(c->p_inbox, cap, world, pass,
            c->x_table.as<uint32_t>(), tsize - 1, d_timeout);
    XEV(); // 1: receive + insert (waits for the lower neighbour)
    PublishArgs pa{};
    pa.peers = peers;
    pa.rank = rank;
    pa.world = world;
    pa.cap = cap;
    pa.pass = pass;
    pa.local_ids = local_ids;
    pa.n_up = d_up;
    pa.n_faces = &dctr->tot.n_faces;
    pa.n_fv = &dctr->tot.n_fv;
    pa.n_zero = &dctr->n_zero;
    pa.known_degenerate = known_degenerate ? 1u : 0u;
    pa.run_overflow = &dctr->overflow;
    pa.gen_err = reinterpret_cast<const unsigned*>(&dctr->gen.err);
    pa.gen_arena_overflow = &dctr->gen.arena_overflow;
    pa.n_bnd_faces = &dctr->gen.n_bnd_faces;
    pa.n_bad = d_bad;
    // persistent blocks take tile tickets until the device-side vertex count is covered
    const uint32_t ms_blocks = std::max<uint32_t>(1, std::min<uint32_t>((c->h_unique + 1023) / 1024, (uint32_t)sm * 4));
    px_mark_scan_publish_kernel<<<ms_blocks, 256, 0, s>>>(c->v_key.as<uint4>(), c->v_size.as<uint8_t>(), d_nv, low, cap,
        c->x_table.as<uint32_t>(), tsize - 1, c->x_has_low ? 1 : 0, c->own_idx.as<uint32_t>(),
        c->x_status.as<unsigned long long>(), d_tile, d_nown, d_done2, pa);
    XEV(); // 2: mark + scan + publish the record
    px_finish_kernel<<<gv, 256, 0, s>>>(c->p_inbox, rank, world, cap, pass, d_voff, d_foff, small, d_nv,
        c->v_key.as<uint4>(), c->own_idx.as<uint32_t>(), c->x_table.as<uint32_t>(), tsize - 1, c->gid.as<uint32_t>(),
        c->v_tet.as<uint32_t>(), c->v_local.as<uint8_t>(), c->v_size.as<uint8_t>(), c->v_simplex.as<uint4>(),
        c->v_funcs.as<uint4>(), c->v_xyz.as<double>(), c->o_tet.as<uint32_t>(), c->o_local.as<uint8_t>(),
        c->o_size.as<uint8_t>(), c->o_simplex.as<uint4>(), c->o_funcs.as<uint4>(), c->o_xyz.as<double>(),
        c->o_key.as<uint4>());
    XEV(); // 3: offsets (waits for every rank's record) + global ids + owned vertices
    c->x_parts = xe;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev_join, s));
    c->launches += c->x_has_low ? 4 : 3;
    return RIN_OK;
}

// joined behind the face kernel: global vertex ids into the face vertex lists, rebased face offsets
int enqueue_fused_apply(rin_ctx* c, Counters* dctr)
{
    cudaStream_t s = c->stream;
    const int rank = c->x_rank, world = c->x_world;
    uint32_t* small = c->x_small.as<uint32_t>();
    uint32_t* d_foff = small + 64 + (world + 1);
    CK(cudaStreamWaitEvent(s, c->ev_join, 0));
    fx_apply_kernel<<<c->sm_count * 4, 256, 0, s>>>(small, c->f_verts.as<uint32_t>(), c->gid.as<uint32_t>(),
        &dctr->tot.n_fv, c->f_off.as<uint32_t>(), c->f_toff.as<uint32_t>(), &dctr->tot.n_faces,
        d_foff + (world + 1) + rank, d_foff + 2 * (world + 1) + rank);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_xsmall, small, (64 + 4 * (size_t)(world + 1)) * 4, cudaMemcpyDeviceToHost, s));
    ++c->launches;
    return RIN_OK;
}

// after the synchronisation, every rank alike: 0 = done, 1 = repeat the pass, X_NOT_FUSED = use the two-call path
int finish_fused_exchange(rin_ctx* c, bool my_run_ok)
{
    const uint32_t* hs = c->h_xsmall;
    const int rank = c->x_rank, world = c->x_world;
    if (hs[10]) return fail(RIN_ERR_STATE, "peer exchange: a peer's message did not arrive within the time limit");
    if (hs[6]) return X_NOT_FUSED;
    if (hs[3]) { // a message outgrew the capacity
        c->x_cap = (hs[3] + hs[3] / 4 + 1024 + 3u) & ~3u;
        return 1;
    }
    if (hs[7] || !my_run_ok) return 1;
    if (hs[5]) return fail(RIN_ERR_STATE, "exchange: " + std::to_string(hs[5]) + " shared vertices have no owner");
    std::swap(c->v_tet, c->o_tet);
    std::swap(c->v_local, c->o_local);
    std::swap(c->v_size, c->o_size);
    std::swap(c->v_simplex, c->o_simplex);
    std::swap(c->v_funcs, c->o_funcs);
    std::swap(c->v_xyz, c->o_xyz);
    std::swap(c->v_key, c->o_key);
    c->n_own = hs[2];
    for (int q = 0; q < 4; ++q) {
        c->x_offsets[2 * q] = hs[64 + q * (world + 1) + rank];
        c->x_offsets[2 * q + 1] = hs[64 + q * (world + 1) + world];
    }
    return RIN_OK;
}
} // namespace
} // extern "C++"

// rin_run + rin_exchange_nccl as ONE call with ONE host synchronisation: the exchange kernels and the NCCL calls are
// enqueued behind the run's kernels (their launch cost hides behind the run) and read every count from device
// memory.  Falls back to the two calls whenever the fused path does not apply (first pass over new inputs, material
// interface, unstructured neighbourhoods, empty ranks, degenerate inputs); every rank takes the same branch.
int rin_run_exchange(rin_ctx* c, int mode, uint32_t flags, uint64_t* vert_offset, uint64_t* n_verts_total,
    uint64_t* face_offset, uint64_t* n_faces_total)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (!c->nccl_comm) return fail(RIN_ERR_STATE, "rin_run_exchange: call rin_nccl_init first");
    const bool can = mode == RIN_MODE_IA && c->x_window && c->x_fusable && !c->fuse_disabled && c->ghost_lo == 0 &&
                     c->ghost_hi == 0 && c->h_act != 0 && c->x_cap != 0 && c->t_count != 0;
    if (can && !c->p_failed && (!c->p_ready || c->x_cap > c->p_cap)) {
        int rc = peer_setup(c); // collective; on any failure every rank keeps the NCCL data path
        if (rc) return rc;
    }
    if (can) {
        c->fuse = true;
        int rc = rin_run(c, mode, flags);
        c->fuse = false;
        if (rc == RIN_OK) {
            if (vert_offset) *vert_offset = c->x_offsets[0];
            if (n_verts_total) *n_verts_total = c->x_offsets[1];
            if (face_offset) *face_offset = c->x_offsets[2];
            if (n_faces_total) *n_faces_total = c->x_offsets[3];
            return RIN_OK;
        }
        if (rc != X_NOT_FUSED) return rc;
        c->fuse_disabled = true; // degenerate inputs (every rank saw the flag): ghost tets, two calls
    }
    int rc = rin_run(c, mode, flags);
    if (rc) return rc;
    return rin_exchange_nccl(c, vert_offset, n_verts_total, face_offset, n_faces_total);
}

// device time (ms) of the exchange chain of the last fused rin_run_exchange: from the end of the vertex kernel to
// the owned-vertex compaction, including the wait for the neighbours (0 when the pass was not fused)
int rin_get_exchange_time(const rin_ctx* c, float* ms)
{
    if (!c || !ms) return fail(RIN_ERR_ARG, "null argument");
    *ms = c->x_chain_ms;
    return RIN_OK;
}

// per-kernel device times (ms) of the peer exchange chain of the last fused pass with stage timing on:
// {send keys, receive + insert, mark + scan + publish record, offsets + global ids + owned vertices}
int rin_get_exchange_parts(const rin_ctx* c, float ms[10])
{
    if (!c || !ms) return fail(RIN_ERR_ARG, "null argument");
    for (int i = 0; i < 10; ++i) ms[i] = c->x_part_ms[i];
    return RIN_OK;
}

int rin_get_exchange_offsets(const rin_ctx* c, uint64_t out[8])
{
    if (!c || !out) return fail(RIN_ERR_ARG, "null argument");
    if (!c->finalized) return fail(RIN_ERR_STATE, "rin_get_exchange_offsets: no finished exchange");
    for (int q = 0; q < 8; ++q) out[q] = c->x_offsets[q];
    return RIN_OK;
}

// introspection for tests: host copy of the IA tables
int rin_debug_ia_tables(rin_ctx* c, const uint16_t** lut1, const uint16_t** lut2, const uint8_t** blob,
    uint32_t* blob_bytes)
{
    if (!c) return fail(RIN_ERR_ARG, "null ctx");
    if (!c->lut_ia.built) {
        CK(cudaSetDevice(c->device));
        int rc = build_ia_tables(c);
        if (rc) return rc;
    }
    *lut1 = c->lut_ia.h_lut1.data();
    *lut2 = c->lut_ia.h_lut2.data();
    *blob = c->lut_ia.h_blob.data();
    *blob_bytes = c->lut_ia.blob_bytes;
    return RIN_OK;
}

} // extern "C"

namespace {

// ------------------------------------------------------------------------------------------------
// One implicit-arrangement pass over tets [t_first, t_first + t_count): eight kernel launches
// (pipeline_ia.cuh) and ONE host synchronisation when the buffer sizes are known from the previous
// pass; the first pass over new inputs synchronises once more after the tile scan to size them.
// ------------------------------------------------------------------------------------------------
template <int W>
int run_ia_w(rin_ctx* c, uint32_t flags)
{
    const uint32_t V = (uint32_t)c->V, VS = c->VS, F = c->F;
    const uint32_t T = (uint32_t)c->t_count, t_first = (uint32_t)c->t_first;
    const int use_lookup = (flags & RIN_FLAG_USE_LOOKUP) ? 1 : 0;
    const int use_secondary = (flags & RIN_FLAG_USE_SECONDARY_LOOKUP) ? 1 : 0;
    const int negate = (flags & RIN_FLAG_NEGATE) ? 1 : 0;
    cudaStream_t s = c->stream;
    const int sm = c->sm_count;
    const bool grid = c->grid_R != 0;
    const bool pack = (F <= 16);
    c->launches = 0;

    CK(c->counters.ensure(sizeof(Counters)));
    Counters* dctr = c->counters.as<Counters>();
    Counters* hp = static_cast<Counters*>(c->h_pinned);

    rin_counts& n = c->counts;
    n = rin_counts{};
    n.num_pts = c->V;
    n.num_tets = c->t_count;
    n.num_funcs = F;
    c->ia_bndry_faces = false;
    c->n_local_verts = c->n_own = 0;
    if (T == 0) { // empty tet range: a valid empty result
        CK(c->f_off.ensure(4));
        CK(c->f_toff.ensure(4));
        CK(cudaMemsetAsync(c->f_off.p, 0, 4, s));
        CK(cudaMemsetAsync(c->f_toff.p, 0, 4, s));
        CK(cudaStreamSynchronize(s));
        for (auto& x : c->stage_ms) x = 0;
        c->kernel_ms[0] = c->kernel_ms[1] = c->total_ms = 0;
        return RIN_OK;
    }

    // ---- buffers whose size depends on the inputs only
    CK(c->vals.ensure((size_t)VS * F * 8));
    if (pack)
        CK(c->vmask16.ensure((size_t)V * 4));
    else
        CK(c->vmask.ensure((size_t)VS * W * 8));
    uint32_t* vm16 = pack ? c->vmask16.as<uint32_t>() : nullptr;
    const uint32_t tile_slots = (uint32_t)filter_tile_slots(W, grid);
    const uint32_t tile_units = (uint32_t)filter_rounds(W, grid) * 256u;
    const uint32_t c_first = grid ? t_first / 5 : 0;
    const uint32_t n_units = grid ? ((t_first + T - 1) / 5 - c_first + 1) : T;
    const uint32_t n_tiles = (n_units + tile_units - 1) / tile_units;
    const size_t tl_stride = (size_t)n_tiles * tile_slots;
    CK(c->tl_tet.ensure(tl_stride * 4));
    CK(c->tl_mask.ensure(tl_stride * 4 * W));
    CK(c->tl_ref.ensure(tl_stride * 4));
    CK(c->tile_tot.ensure((size_t)n_tiles * sizeof(TileTot)));
    CK(c->tile_pre.ensure((size_t)(n_tiles + 1) * sizeof(TileTot)));
    const uint32_t last_mask = (F % 32) ? ((1u << (F % 32)) - 1u) : 0xffffffffu;
    const size_t small_smem = GEN_SMALL_WARPS * sizeof(SmallSlot);
    const size_t mid_smem = GEN_MID_WARPS * sizeof(MidSlot);
    const int mid_per_sm = (int)std::max<size_t>(1, c->smem_per_sm / (mid_smem + 1024));
    if (!(c->attr_set & (1u << W))) { // once per context and mask width (these calls sit in front of the first launch)
        if (mid_smem > 48 * 1024)
            CK(cudaFuncSetAttribute(general_ia_mid_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mid_smem));
        CK(cudaFuncSetAttribute(general_ia_mid_kernel<W>, cudaFuncAttributePreferredSharedMemoryCarveout,
            (int)cudaSharedmemCarveoutMaxShared));
        c->attr_set |= 1u << W;
    }
    if (c->arena.cap == 0) CK(c->arena.ensure(1u << 20));

    Counters h{};
    bool evaluated = false;
    unsigned long long n_zero = 0; // counted by the evaluation kernel, which runs in the first attempt only
    for (int attempt = 0;; ++attempt) {
        if (attempt > 5) return fail(RIN_ERR_STATE, "implicit arrangement: buffer sizing did not converge");
        const bool sizing = (c->h_act == 0); // sizes unknown: one extra synchronisation after the tile scan
        const uint32_t list_cap = std::max<uint32_t>(c->h_list, 4096);
        CK(c->general_list.ensure((size_t)list_cap * 4));
        CK(c->big_list.ensure((size_t)list_cap * 8)); // [big | small-tier overflow]
        uint32_t act_cap = 0, cand_cap = 0, face_cap = 0, fv_cap = 0, tsize = 0;
        auto size_outputs = [&]() -> int {
            act_cap = c->h_act;
            cand_cap = c->h_cand;
            face_cap = c->h_face;
            fv_cap = c->h_fv;
            c->act_cap = std::max(c->act_cap, act_cap);
            CK(c->act_tet.ensure((size_t)c->act_cap * 4));
            CK(c->act_mask.ensure((size_t)c->act_cap * 4 * W));
            CK(c->rec_ref.ensure((size_t)c->act_cap * 4));
            CK(c->offs.ensure((size_t)c->act_cap * 16));
            CK(c->status.ensure(((size_t)cand_cap / RV_TILE + 2) * 8 + 64));
            CK(c->cand_src.ensure((size_t)cand_cap * 4));
            CK(c->cand_key.ensure((size_t)cand_cap * 16));
            CK(c->slot_of.ensure((size_t)cand_cap * 4));
            // slots: twice the unique keys of the previous pass (a full table is detected and the pass repeated),
            // or twice the candidates when nothing is known
            tsize = 1024;
            while (tsize < 2ull * (c->h_unique ? c->h_unique : cand_cap)) tsize <<= 1;
            CK(c->table.ensure((size_t)tsize * 4));
            CK(c->v_tet.ensure((size_t)cand_cap * 4));
            CK(c->v_local.ensure(cand_cap));
            CK(c->v_size.ensure(cand_cap));
            CK(c->v_simplex.ensure((size_t)cand_cap * 16));
            CK(c->v_funcs.ensure((size_t)cand_cap * 16));
            CK(c->v_xyz.ensure((size_t)cand_cap * 24));
            CK(c->v_key.ensure((size_t)cand_cap * 16));
            CK(c->f_off.ensure((size_t)(face_cap + 1) * 4));
            CK(c->f_toff.ensure((size_t)(face_cap + 1) * 4));
            CK(c->f_tets.ensure((size_t)face_cap * 8));
            CK(c->f_funcs.ensure((size_t)face_cap * 8));
            CK(c->f_verts.ensure((size_t)fv_cap * 4));
            return RIN_OK;
        };
        if (!sizing) {
            int rc = size_outputs();
            if (rc) return rc;
        }

        CK(cudaMemsetAsync(dctr, 0, sizeof(Counters), s));

        // ---- K1: values + sign masks (only the vertex range the tet range touches)
        EVREC(c->ev[ST_EVAL]);
        EVREC(c->kev[0]);
        if (!evaluated) {
            const uint32_t vf = c->v_count ? c->v_first : 0, vc = c->v_count ? c->v_count : V;
            if (c->have_funcs) {
                const size_t smem = F * sizeof(rin_func_desc);
                const int g = grid_for((vc + EV_VPT - 1) / EV_VPT, 256, sm, 8);
                if (grid)
                    eval_kernel<true><<<g, 256, smem, s>>>(nullptr, c->axes.as<double>(), c->grid_R + 1,
                        make_fastdiv(c->grid_R + 1), vf, vc, VS,
                        c->funcs.as<rin_func_desc>(), F, negate, c->vals.as<double>(), c->vmask.as<uint2>(), vm16,
                        &dctr->n_zero);
                else
                    eval_kernel<false><<<g, 256, smem, s>>>(c->pts.as<double>(), nullptr, 0, FastDiv{}, vf, vc, VS,
                        c->funcs.as<rin_func_desc>(), F, negate, c->vals.as<double>(), c->vmask.as<uint2>(), vm16,
                        &dctr->n_zero);
            } else {
                ingest_values_kernel<<<grid_for(vc, 256, sm, 8), 256, 0, s>>>(c->rowmajor.as<double>(), vf, vc, VS, F,
                    negate, c->vals.as<double>(), c->vmask.as<uint2>(), vm16, &dctr->n_zero);
            }
            CK(cudaGetLastError());
            ++c->launches;
        }
        EVREC(c->kev[1]);

        // ---- K2: filter + dispatch
        EVREC(c->ev[ST_FILTER]);
        FilterArgs fa{};
        fa.tets = grid ? nullptr : c->tets.as<uint4>();
        fa.R = c->grid_R;
        fa.dR = make_fastdiv(c->grid_R);
        fa.t_first = t_first;
        fa.t_count = T;
        fa.c_first = c_first;
        fa.n_units = n_units;
        fa.vmask = c->vmask.as<uint2>();
        fa.vmask16 = vm16;
        fa.VS = VS;
        fa.last_mask = last_mask;
        fa.vals = c->vals.as<double>();
        fa.lut1 = c->lut_ia.lut1.as<uint16_t>();
        fa.lut2 = c->lut_ia.lut2.as<uint16_t>();
        fa.blob32 = c->lut_ia.blob.as<uint32_t>();
        fa.arena32 = c->arena.as<uint32_t>();
        fa.use_lookup = use_lookup;
        fa.use_secondary = use_secondary;
        fa.tl_tet = c->tl_tet.as<uint32_t>();
        fa.tl_mask = c->tl_mask.as<uint32_t>();
        fa.tl_ref = c->tl_ref.as<uint32_t>();
        fa.tl_stride = tl_stride;
        fa.tile_tot = c->tile_tot.as<TileTot>();
        fa.small_list = c->general_list.as<uint32_t>();
        fa.big_list = c->big_list.as<uint32_t>();
        fa.list_cap = list_cap;
        fa.fc = &dctr->filt;
        fa.gc = &dctr->gen;
        fa.n_exact = &dctr->n_exact_classify;
        fa.overflow = &dctr->overflow;
        EVREC(c->kev[2]);
        if (grid) {
            if (pack && W == 1)
                filter_classify_kernel<1, true, true><<<n_tiles, 256, 0, s>>>(fa);
            else
                filter_classify_kernel<W, false, true><<<n_tiles, 256, 0, s>>>(fa);
        } else {
            if (pack && W == 1)
                filter_classify_kernel<1, true, false><<<n_tiles, 256, 0, s>>>(fa);
            else
                filter_classify_kernel<W, false, false><<<n_tiles, 256, 0, s>>>(fa);
        }
        EVREC(c->kev[3]);
        CK(cudaGetLastError());
        ++c->launches;

        // ---- K4: general tiers; the last block of the big tier scans the tile totals
        EVREC(c->ev[ST_CLASSIFY]);
        EVREC(c->ev[ST_GENERAL]);
        const uint32_t acap = (uint32_t)std::min<size_t>(c->arena.cap, 0xfffffff0u);
        const uint32_t est = std::min<uint32_t>(list_cap, c->h_list ? c->h_list : list_cap);
        const int small_blocks = (int)std::max<uint32_t>(
            1, std::min<uint32_t>((est + GEN_SMALL_WARPS - 1) / GEN_SMALL_WARPS, (uint32_t)sm * 14));
        const int mid_blocks = (int)std::max<uint32_t>(
            1, std::min<uint32_t>((est + GEN_MID_WARPS - 1) / GEN_MID_WARPS, (uint32_t)(sm * mid_per_sm)));
        uint32_t* lists = c->big_list.as<uint32_t>();
        TileScanArgs sa{};
        sa.tot = fa.tile_tot;
        sa.n_tiles = n_tiles;
        sa.off = c->tile_pre.as<TileTot>();
        sa.totals = &dctr->tot;
        sa.act_cap = sizing ? 0xffffffffu : act_cap;
        sa.cand_cap = sizing ? 0xffffffffu : cand_cap;
        sa.face_cap = sizing ? 0xffffffffu : face_cap;
        sa.fv_cap = sizing ? 0xffffffffu : fv_cap;
        sa.overflow = &dctr->overflow;
        general_ia_small_kernel<W><<<small_blocks, GEN_SMALL_WARPS * 32, small_smem, s>>>(c->tets.as<uint4>(), fa.tl_tet,
            fa.tl_mask, (uint32_t)tl_stride, fa.small_list, lists + list_cap,
            c->lut_ia.built ? c->lut_ia.cx2.as<IAComplex<IACapsSmall>>() : nullptr, c->lut_ia.lut2cx.as<uint16_t>(),
            fa.vals, VS, c->arena.as<uint8_t>(), acap, fa.tl_ref, &dctr->gen, fa.tile_tot, tile_slots, list_cap);
        // The mid tier's launch costs ~10 us even when it has nothing to do (its shared-memory carve-out makes the SMs
        // reconfigure twice).  When the previous pass had no tet for it, it is left out; the read-back tells whether
        // this pass had one after all, and then the pass is repeated with it (never in a fused run + exchange, where
        // every rank must take the same branch).
        const bool skip_mid = c->skip_mid && !c->fuse && !sizing;
        if (!skip_mid)
            general_ia_mid_kernel<W><<<mid_blocks, GEN_MID_WARPS * 32, mid_smem, s>>>(c->tets.as<uint4>(), fa.tl_tet,
                fa.tl_mask, (uint32_t)tl_stride, lists, lists + list_cap, nullptr, fa.vals, VS, c->arena.as<uint8_t>(),
                acap, fa.tl_ref, &dctr->gen, fa.tile_tot, tile_slots, list_cap, TileScanArgs{});
        // exclusive scan of the per-tile totals: one block of 32 warps (the last block of the mid tier has four warps
        // only: 13 us at 2048 tiles, 80 us at 16384)
        scan_tiles5_kernel<<<1, 1024, 0, s>>>(sa);
        CK(cudaGetLastError());
        c->launches += skip_mid ? 2 : 3;
        EVREC(c->ev[ST_SCAN]);

        auto check_general = [&](const Counters& hc, bool& again) -> int {
            again = false;
            if (hc.gen.err)
                return fail(hc.gen.err, "per-tet arrangement failed in tet " + std::to_string(hc.gen.err_tet) +
                                            (hc.gen.err == RIN_ERR_CAPACITY ? " (complex exceeds kernel capacity)"
                                                                           : " (degenerate input plane)"));
            if (hc.gen.arena_overflow) {
                CK(c->arena.ensure((size_t)hc.gen.arena_top + hc.gen.arena_top / 8 + 4096));
                again = true;
            }
            if (hc.overflow & OVF_LIST) {
                const uint32_t need = std::max({hc.gen.n_small, hc.gen.n_big, hc.gen.n_ovf});
                c->h_list = need + need / 8 + 1024;
                again = true;
            }
            return RIN_OK;
        };
        if (sizing) {
            CK(cudaMemcpyAsync(hp, dctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            h = *hp;
            if (!evaluated) n_zero = h.n_zero;
            evaluated = true;
            bool again;
            int rc = check_general(h, again);
            if (rc) return rc;
            if (again) continue;
            c->h_act = h.tot.n_active + h.tot.n_active / 8 + 1024;
            c->h_cand = h.tot.n_cand + h.tot.n_cand / 8 + 1024;
            c->h_face = h.tot.n_faces + h.tot.n_faces / 8 + 1024;
            c->h_fv = h.tot.n_fv + h.tot.n_fv / 8 + 1024;
            rc = size_outputs();
            if (rc) return rc;
        }

        // ---- K5 + K6: ordered active list, candidates, hash-min insertion
        EVREC(c->ev[ST_EMIT]);
        CK(cudaMemsetAsync(c->table.p, 0xff, (size_t)tsize * 4, s));
        CK(cudaMemsetAsync(c->status.p, 0, ((size_t)cand_cap / RV_TILE + 2) * 8, s));
        EmitArgs ea{};
        ea.tets = c->tets.as<uint4>();
        ea.R = c->grid_R;
        ea.dR = fa.dR;
        ea.tile_ticket = &dctr->scan.tile_counter;
        ea.tl_tet = fa.tl_tet;
        ea.tl_mask = fa.tl_mask;
        ea.tl_ref = fa.tl_ref;
        ea.tl_stride = tl_stride;
        ea.tile_slots = tile_slots;
        ea.tile_tot = fa.tile_tot;
        ea.tile_off = c->tile_pre.as<TileTot>();
        ea.n_tiles = n_tiles;
        ea.blob32 = fa.blob32;
        ea.arena32 = c->arena.as<uint32_t>();
        ea.act_tet = c->act_tet.as<uint32_t>();
        ea.act_mask = c->act_mask.as<uint32_t>();
        ea.act_cap = c->act_cap;
        ea.rec_ref = c->rec_ref.as<uint32_t>();
        ea.offs = c->offs.as<uint4>();
        ea.cand_key = c->cand_key.as<uint4>();
        ea.cand_src = c->cand_src.as<uint32_t>();
        ea.overflow = &dctr->overflow;
        emit_kernel<W><<<(int)std::min<uint32_t>(n_tiles, (uint32_t)sm * 8), 256, 0, s>>>(ea);
        const uint32_t a_est = sizing ? h.tot.n_active : act_cap;
        const uint32_t c_est = sizing ? h.tot.n_cand : cand_cap;
        InsertArgs ia{};
        ia.cand_key = ea.cand_key;
        ia.totals = &dctr->tot;
        ia.table = c->table.as<uint32_t>();
        ia.table_mask = tsize - 1;
        ia.slot_of = c->slot_of.as<uint32_t>();
        ia.overflow = &dctr->overflow;
        EVREC(c->ev[ST_DEDUP]);
        insert_kernel<<<grid_for((c_est + INS_BATCH - 1) / INS_BATCH, 256, sm, 8), 256, 0, s>>>(ia);
        CK(cudaGetLastError());
        c->launches += 2;

        // ---- K6 + K7: ranking, vertex records, coordinates
        EVREC(c->ev[ST_VERTS]);
        RankArgs ra{};
        ra.tets = ea.tets;
        ra.R = ea.R;
        ra.dR = ea.dR;
        ra.act_tet = ea.act_tet;
        ra.act_mask = ea.act_mask;
        ra.act_cap = c->act_cap;
        ra.rec_ref = ea.rec_ref;
        ra.offs = ea.offs;
        ra.blob32 = ea.blob32;
        ra.arena32 = ea.arena32;
        ra.totals = &dctr->tot;
        ra.cand_src = ea.cand_src;
        ra.table = ia.table;
        ra.slot_of = ia.slot_of;
        ra.vals = fa.vals;
        ra.VS = VS;
        ra.pts = c->pts.as<double>();
        ra.v_tet = c->v_tet.as<uint32_t>();
        ra.v_local = c->v_local.as<uint8_t>();
        ra.v_size = c->v_size.as<uint8_t>();
        ra.v_simplex = c->v_simplex.as<uint4>();
        ra.v_funcs = c->v_funcs.as<uint4>();
        ra.v_xyz = c->v_xyz.as<double>();
        ra.v_key = c->v_key.as<uint4>();
        ra.status = c->status.as<unsigned long long>();
        ra.tile_counter = &dctr->rank_tile;
        ra.n_unique = &dctr->n_unique;
        ra.overflow = &dctr->overflow;
        rank_verts_kernel<W><<<(int)std::max<uint32_t>(1, std::min<uint32_t>((c_est + RV_TILE - 1) / RV_TILE,
                                   (uint32_t)sm * 6)), 256, 0, s>>>(ra);
        CK(cudaGetLastError());
        ++c->launches;
        if (c->fuse) { // rin_run_exchange: the vertex exchange starts here, on a second stream beside the face kernel
            const bool peer = c->p_ready && c->x_cap <= c->p_cap; // the same on every rank
            int rc = peer ? enqueue_peer_exchange(c, dctr, evaluated && n_zero != 0, cand_cap)
                          : enqueue_fused_exchange(c, dctr, evaluated && n_zero != 0, cand_cap);
            if (rc) return rc;
        }

        // ---- faces
        EVREC(c->ev[ST_FACES]);
        FaceArgs fg{};
        fg.act_tet = ea.act_tet;
        fg.act_mask = ea.act_mask;
        fg.act_cap = c->act_cap;
        fg.rec_ref = ea.rec_ref;
        fg.offs = ea.offs;
        fg.blob32 = ea.blob32;
        fg.arena32 = ea.arena32;
        fg.totals = &dctr->tot;
        fg.table = ia.table;
        fg.slot_of = ia.slot_of;
        fg.f_off = c->f_off.as<uint32_t>();
        fg.f_verts = c->f_verts.as<uint32_t>();
        fg.f_toff = c->f_toff.as<uint32_t>();
        fg.f_tets = c->f_tets.as<uint32_t>();
        fg.f_funcs = c->f_funcs.as<uint32_t>();
        fg.overflow = &dctr->overflow;
        // fused run + exchange: many short blocks instead of persistent ones, so that the high-priority exchange
        // stream finds free SM slots while the face kernel runs
        const int face_grid = c->fuse ? grid_for(a_est, 256, sm, 64) : grid_for(a_est, 256, sm, 8);
        faces_kernel<W, false><<<face_grid, 256, 0, s>>>(fg);
        CK(cudaGetLastError());
        ++c->launches;
        EVREC(c->ev[ST_COUNT]);
        if (c->fuse) { // ... and is joined here, same synchronisation as the run's
            int rc = enqueue_fused_apply(c, dctr);
            if (rc) return rc;
        }

        // ---- the one read-back
        CK(cudaMemcpyAsync(hp, dctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        h = *hp;
        if (!evaluated) n_zero = h.n_zero;
        evaluated = true;
        bool again;
        int rc = check_general(h, again);
        if (rc) return rc;
        if (skip_mid && (h.gen.n_big || h.gen.n_ovf)) { // the mid tier had work after all: once more, with it
            c->skip_mid = false;
            continue;
        }
        const bool my_ok = !(again || h.overflow);
        int fr = RIN_OK;
        if (c->fuse) { // the same decision on every rank (it comes from the gathered records)
            fr = finish_fused_exchange(c, my_ok);
            if (fr == X_NOT_FUSED || fr < 0) return fr;
            cudaEventElapsedTime(&c->x_chain_ms, c->ev_fork, c->ev_join);
            for (int i = 0; i < 10; ++i) c->x_part_ms[i] = 0;
            if (c->stage_timing)
                for (int i = 0; i < c->x_parts; ++i)
                    cudaEventElapsedTime(&c->x_part_ms[i], i ? c->ev_x[i - 1] : c->ev_fork, c->ev_x[i]);
        }
        if (!my_ok) {
            c->h_act = 0; // size the next attempt exactly
            c->h_unique = 0;
            continue;
        }
        if (fr == 1) continue; // another rank has to repeat its pass

        uint32_t NF = h.tot.n_faces, NFVout = h.tot.n_fv, NFT = h.tot.n_faces;
        if (h.gen.n_bnd_faces) {
            // degenerate input: iso-faces on tet boundaries are shared between two tets
            // (src/extract_mesh.cpp:240-253); the vertex table stays untouched (rin_tet_maps reads it)
            const uint32_t NFc = h.tot.n_faces, NFV = h.tot.n_fv;
            uint32_t t2 = 1024;
            while (t2 < 2 * NFc) t2 <<= 1;
            CK(c->ftable.ensure((size_t)t2 * 4));
            CK(c->bids.ensure((size_t)NFc * 4));
            CK(c->face_hdr.ensure((size_t)NFc * 16));
            CK(c->tmp_fverts.ensure((size_t)std::max(NFV, 1u) * 4));
            CK(c->bfkeys.ensure((size_t)NFc * 16));
            CK(c->frep.ensure((size_t)NFc * 4));
            CK(c->fdup.ensure((size_t)NFc * 8 + 16)); // ndup | cursor | totals
            CK(c->fpos.ensure((size_t)NFc * 16));
            uint32_t* ndup = c->fdup.as<uint32_t>();
            uint32_t* cursor = ndup + NFc;
            uint32_t* totals = cursor + NFc;
            CK(cudaMemsetAsync(c->ftable.p, 0xff, (size_t)t2 * 4, s));
            CK(cudaMemsetAsync(c->fdup.p, 0, (size_t)NFc * 8 + 16, s));
            FaceArgs fh = fg;
            fh.f_verts = c->tmp_fverts.as<uint32_t>();
            fh.face_hdr = c->face_hdr.as<uint4>();
            faces_kernel<W, true><<<face_grid, 256, 0, s>>>(fh);
            const int g = grid_for(NFc, 256, sm, 8);
            bface_keys_kernel<<<g, 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc, c->tmp_fverts.as<uint32_t>(),
                c->bfkeys.as<uint4>());
            bface_insert_kernel<<<g, 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc, c->bfkeys.as<uint4>(),
                c->ftable.as<uint32_t>(), t2 - 1, c->bids.as<uint32_t>());
            bface_reps_kernel<<<g, 256, 0, s>>>(c->ftable.as<uint32_t>(), c->bids.as<uint32_t>(), NFc,
                c->frep.as<uint32_t>(), ndup);
            bface_scan_kernel<<<1, 1024, 0, s>>>(c->face_hdr.as<uint4>(), NFc, c->frep.as<uint32_t>(), ndup,
                c->fpos.as<uint4>(), totals);
            bface_write_kernel<<<grid_for((uint64_t)NFc + 1, 256, sm, 8), 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc,
                c->tmp_fverts.as<uint32_t>(), c->frep.as<uint32_t>(), c->fpos.as<uint4>(), cursor, totals,
                c->f_off.as<uint32_t>(), c->f_verts.as<uint32_t>(), c->f_toff.as<uint32_t>(),
                c->f_tets.as<uint32_t>(), c->f_funcs.as<uint32_t>(), 0);
            CK(cudaGetLastError());
            uint32_t ht[3];
            CK(cudaMemcpyAsync(ht, totals, 12, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            NF = ht[0];
            NFVout = ht[1];
            NFT = ht[2];
            bface_sort_pairs_kernel<<<grid_for(NF, 256, sm, 8), 256, 0, s>>>(NF, c->f_toff.as<uint32_t>(),
                c->f_tets.as<uint32_t>());
            CK(cudaGetLastError());
            c->launches += 7;
            EVREC(c->ev[ST_COUNT]);
            CK(cudaStreamSynchronize(s));
        }
        if (c->stage_timing) {
            for (int i = 0; i < ST_COUNT; ++i) CK(cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]));
            CK(cudaEventElapsedTime(&c->kernel_ms[0], c->kev[0], c->kev[1]));
            CK(cudaEventElapsedTime(&c->kernel_ms[1], c->kev[2], c->kev[3]));
        } else {
            for (auto& x : c->stage_ms) x = 0;
            c->kernel_ms[0] = c->kernel_ms[1] = 0;
        }
        CK(cudaEventElapsedTime(&c->total_ms, c->ev[0], c->ev[ST_COUNT]));

        // learnt sizes for the next pass over these inputs
        c->h_act = h.tot.n_active + h.tot.n_active / 8 + 1024;
        c->h_cand = h.tot.n_cand + h.tot.n_cand / 8 + 1024;
        c->h_face = h.tot.n_faces + h.tot.n_faces / 8 + 1024;
        c->h_fv = h.tot.n_fv + h.tot.n_fv / 8 + 1024;
        {
            const uint32_t need = std::max({h.gen.n_small, h.gen.n_big, h.gen.n_ovf});
            c->h_list = need + need / 8 + 1024;
        }
        c->table_size = tsize;
        c->skip_mid = h.gen.n_big == 0 && h.gen.n_ovf == 0;
        c->h_unique = h.n_unique + h.n_unique / 8 + 1024;
        c->ia_bndry_faces = h.gen.n_bnd_faces != 0;
        const uint32_t NV = h.tot.n_cand ? h.n_unique : 0;
        c->n_local_verts = NV;
        c->n_own = NV;
        n.num_degenerate_vertex = n_zero;
        n.num_intersecting_tet = h.tot.n_active;
        n.num_k1 = h.filt.n_k1;
        n.num_k2 = h.filt.n_k2;
        n.num_kmore = h.filt.n_kmore;
        n.num_active_funcs = h.tot.n_funcs;
        n.num_general_tets = (uint64_t)h.gen.n_small + h.gen.n_big;
        n.num_verts = NV;
        n.num_faces = NF;
        n.num_face_verts = NFVout;
        n.num_face_tets = NFT;
        n.num_exact_fallbacks = (uint64_t)h.gen.n_exact + h.n_exact_classify;
        if (c->fuse) { // the exchange is done: owned vertices only, offsets rebased into the merged mesh
            c->n_own = c->h_xsmall[2];
            n.num_verts = c->n_own;
            c->marked = c->finalized = true;
        }
        return RIN_OK;
    }
}

// ------------------------------------------------------------------------------------------------
// One material-interface pass over tets [t_first, t_first + t_count).
// ------------------------------------------------------------------------------------------------
template <int W>
int run_mi_w(rin_ctx* c, uint32_t flags)
{
    const uint32_t V = c->VS, F = c->F; // V: row stride of vals / vmask
    const uint32_t T = (uint32_t)c->t_count, t_first = (uint32_t)c->t_first;
    if (T == 0) { // empty tet range: a valid empty result
        c->counts = rin_counts{};
        c->counts.num_pts = c->V;
        c->counts.num_funcs = F;
        CK(c->f_off.ensure(4));
        CK(c->f_toff.ensure(4));
        CK(cudaMemsetAsync(c->f_off.p, 0, 4, c->stream));
        CK(cudaMemsetAsync(c->f_toff.p, 0, 4, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->n_local_verts = c->n_own = 0;
        for (auto& x : c->stage_ms) x = 0;
        c->kernel_ms[0] = c->kernel_ms[1] = c->total_ms = 0;
        return RIN_OK;
    }
    const int use_lookup = (flags & RIN_FLAG_USE_LOOKUP) ? 1 : 0;
    const int use_secondary = (flags & RIN_FLAG_USE_SECONDARY_LOOKUP) ? 1 : 0;
    const int negate = (flags & RIN_FLAG_NEGATE) ? 1 : 0;
    cudaStream_t s = c->stream;
    const int sm = c->sm_count;

    CK(c->counters.ensure(sizeof(Counters)));
    Counters* dctr = c->counters.as<Counters>();
    Counters h{};

    // ---- K1: values + sign masks.  The vertex range is restricted to what the tet range can touch
    // only for generated grids by the caller (rin_set_tet_range keeps all V by default).
    CK(c->vals.ensure((size_t)V * F * 8));
    CK(c->vmask.ensure((size_t)V * W * 8));
    CK(cudaMemsetAsync(dctr, 0, sizeof(Counters), s));
    EVREC(c->ev[ST_EVAL]);
    const uint32_t vf = c->v_count ? c->v_first : 0, vc = c->v_count ? c->v_count : (uint32_t)c->V;
    EVREC(c->kev[0]);
    c->launches = 0;
    if (c->have_funcs && F <= (uint32_t)MI_EVAL_MAXF) {
        // evaluation fused with the "highest func" loop: the values never come back from HBM
        const size_t smem = mi_eval_smem(F);
        if (!(c->attr_set & (1u << 16))) {
            CK(cudaFuncSetAttribute(eval_mi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi_eval_smem(MI_EVAL_MAXF)));
            CK(cudaFuncSetAttribute(eval_mi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi_eval_smem(MI_EVAL_MAXF)));
            c->attr_set |= 1u << 16;
        }
        const int g = grid_for((vc + MI_EVAL_VPT - 1) / MI_EVAL_VPT, 256, sm, 8);
        if (c->grid_R)
            eval_mi_kernel<true><<<g, 256, smem, s>>>(nullptr, c->axes.as<double>(), c->grid_R + 1,
                make_fastdiv(c->grid_R + 1), vf, vc, V, c->funcs.as<rin_func_desc>(), F, negate,
                c->vals.as<double>(), c->vmask.as<uint2>(), &dctr->n_zero);
        else
            eval_mi_kernel<false><<<g, 256, smem, s>>>(c->pts.as<double>(), nullptr, 0, FastDiv{}, vf, vc, V,
                c->funcs.as<rin_func_desc>(), F, negate, c->vals.as<double>(), c->vmask.as<uint2>(), &dctr->n_zero);
        ++c->launches;
    } else {
        if (c->have_funcs) {
            const size_t smem = F * sizeof(rin_func_desc);
            const int g = grid_for((vc + EV_VPT - 1) / EV_VPT, 256, sm, 8);
            if (c->grid_R)
                eval_kernel<true><<<g, 256, smem, s>>>(nullptr, c->axes.as<double>(), c->grid_R + 1,
                    make_fastdiv(c->grid_R + 1), vf, vc, V,
                    c->funcs.as<rin_func_desc>(), F, negate, c->vals.as<double>(), c->vmask.as<uint2>(), nullptr,
                    &dctr->n_zero);
            else
                eval_kernel<false><<<g, 256, smem, s>>>(c->pts.as<double>(), nullptr, 0, FastDiv{}, vf, vc, V,
                    c->funcs.as<rin_func_desc>(), F, negate, c->vals.as<double>(), c->vmask.as<uint2>(), nullptr,
                    &dctr->n_zero);
        } else {
            ingest_values_kernel<<<grid_for(vc, 256, sm, 8), 256, 0, s>>>(c->rowmajor.as<double>(), vf, vc, V, F,
                negate, c->vals.as<double>(), c->vmask.as<uint2>(), nullptr, &dctr->n_zero);
        }
        // "highest func": masks of the maximal materials replace the sign masks
        CK(cudaMemsetAsync(&dctr->n_zero, 0, 8, s));
        highest_material_kernel<<<grid_for(vc, 256, sm, 8), 256, 0, s>>>(c->vals.as<double>(), vf, vc, V, F,
            c->vmask.as<uint2>(), &dctr->n_zero);
        c->launches += 2;
    }
    EVREC(c->kev[1]);
    CK(cudaGetLastError());

    // ---- K2: filter (tile-local compaction) + tile scan + ordered gather
    EVREC(c->ev[ST_FILTER]);
    // generated grids: cube-structured filter (the tet index stream is not read)
    const bool mgrid = c->grid_R != 0;
    const uint32_t tile_slots = mgrid ? (uint32_t)MI_GRID_SLOTS : (uint32_t)FILT_TILE;
    const uint32_t c_first = mgrid ? t_first / 5 : 0;
    const uint32_t n_units = mgrid ? ((t_first + T - 1) / 5 - c_first + 1) : T;
    const uint32_t n_tiles = mgrid ? (n_units + MI_GRID_ROUNDS * 256 - 1) / (MI_GRID_ROUNDS * 256)
                                   : (T + FILT_TILE - 1) / FILT_TILE;
    const size_t tl_stride = (size_t)n_tiles * tile_slots;
    CK(c->tl_tet.ensure(tl_stride * 4));
    CK(c->tl_mask.ensure(tl_stride * 4 * W));
    CK(c->tile_cnt.ensure((size_t)n_tiles * 8));
    CK(c->tile_off.ensure((size_t)(n_tiles + 1) * 8));
    EVREC(c->kev[2]);
    if (mgrid)
        filter_mi_grid_kernel<W><<<n_tiles, 256, 0, s>>>(c->grid_R, make_fastdiv(c->grid_R), t_first, T, c_first,
            n_units, c->vmask.as<uint2>(), c->vals.as<double>(), V, F, c->tl_tet.as<uint32_t>(),
            c->tl_mask.as<uint32_t>(), tl_stride, c->tile_cnt.as<uint2>(), &dctr->filt);
    else
        filter_mi_tiles_kernel<W><<<n_tiles, FILT_THREADS, 0, s>>>(c->tets.as<uint4>(), t_first, T,
            c->vmask.as<uint2>(), c->vals.as<double>(), V, F, c->tl_tet.as<uint32_t>(), c->tl_mask.as<uint32_t>(),
            tl_stride, c->tile_cnt.as<uint2>(), &dctr->filt);
    EVREC(c->kev[3]);
    scan_tiles_kernel<<<1, 1024, 0, s>>>(c->tile_cnt.as<uint2>(), n_tiles, c->tile_off.as<uint2>(), &dctr->filt);
    CK(cudaGetLastError());
    c->launches += 2;
    // The number of active tets is read back here only when nothing is known about it; otherwise the kernels up to
    // the count + scan are launched for the capacity learnt from the previous pass and read the count from device
    // memory, and the first read-back of the pass is the merged one behind the count + scan.
    const bool nosync = c->h_act_mi != 0 && c->act_cap >= c->h_act_mi && getenv("RIN_MI_SYNC") == nullptr;
    const unsigned* dA = nosync ? &dctr->filt.n_active : nullptr;
    uint32_t A = 0;  // active tets (exact once read back)
    uint32_t An = 0; // launch bound and list stride until then
    if (!nosync) {
        CK(cudaMemcpyAsync(&h, dctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        A = An = h.filt.n_active;
        c->act_cap = std::max<uint32_t>(c->act_cap, A + A / 8 + 1024);
    } else
        A = An = c->h_act_mi;
    CK(c->act_tet.ensure((size_t)c->act_cap * 4));
    CK(c->act_mask.ensure((size_t)c->act_cap * 4 * W));
    if (An) {
        compact_active_kernel<W><<<grid_for(An, 256, sm, 8), 256, 0, s>>>(c->tl_tet.as<uint32_t>(),
            c->tl_mask.as<uint32_t>(), tl_stride, c->tile_off.as<uint2>(), n_tiles, An, c->act_tet.as<uint32_t>(),
            c->act_mask.as<uint32_t>(), c->act_cap, tile_slots, dA);
        CK(cudaGetLastError());
        ++c->launches;
    }

    // ---- K3: classify
    EVREC(c->ev[ST_CLASSIFY]);
    CK(c->rec_ref.ensure((size_t)std::max(An, 1u) * 4));
    CK(c->general_list.ensure((size_t)std::max(An, 1u) * 4));
    CK(c->big_list.ensure((size_t)std::max(An, 1u) * 8)); // [big | small-tier overflow]
    CK(c->offs.ensure((size_t)std::max(An, 1u) * 16));
    if (An) {
        classify_mi_kernel<W><<<grid_for(An, 256, sm, 4), 256, 0, s>>>(c->tets.as<uint4>(),
            c->act_tet.as<uint32_t>(), c->act_mask.as<uint32_t>(), c->act_cap, An, c->vals.as<double>(),
            c->vmask.as<uint2>(), V, F, c->lut_mi.lut1.as<uint16_t>(), c->lut_mi.lut2.as<uint32_t>(), use_lookup,
            use_secondary, c->rec_ref.as<uint32_t>(),
            c->general_list.as<uint32_t>(), c->big_list.as<uint32_t>(), &dctr->gen, &dctr->n_tie_faces, dA);
        CK(cudaGetLastError());
        ++c->launches;
    }

    // ---- K4 + K5a: general kernels (arena grows on overflow), counts + offsets; ONE read-back for both
    EVREC(c->ev[ST_GENERAL]);
    const uint32_t a_tiles = (An + CS_TILE - 1) / CS_TILE;
    unsigned n_gated = 0;
    if (An) {
        CK(c->status.ensure((size_t)a_tiles * 16 + 64));
        const uint32_t est = nosync ? (use_lookup ? c->h_gen_mi : An)
                                    : (use_lookup ? (h.filt.n_kmore + h.filt.n_k2) : An);
        const int small_blocks = (int)std::max<uint32_t>(
            1, std::min<uint32_t>((est + 64 + GEN_SMALL_WARPS - 1) / GEN_SMALL_WARPS, (uint32_t)sm * 14));
        const size_t small_smem = GEN_SMALL_WARPS * sizeof(MISmallSlot);
        for (int attempt = 0;; ++attempt) {
            if (c->arena.cap == 0) CK(c->arena.ensure(1u << 20));
            const uint32_t acap = (uint32_t)std::min<size_t>(c->arena.cap, 0xfffffff0u);
            const unsigned top0 = 4;
            CK(cudaMemsetAsync(c->arena.p, 0, 4, s));
            CK(cudaMemcpyAsync(&dctr->gen.arena_top, &top0, 4, cudaMemcpyHostToDevice, s));
            // small tier: one tet per warp, complex in shared memory; capacity overflow -> ovf list
            general_mi_small_kernel<W><<<small_blocks, GEN_SMALL_WARPS * 32, small_smem, s>>>(
                c->tets.as<uint4>(), c->act_tet.as<uint32_t>(), c->act_mask.as<uint32_t>(), c->act_cap,
                c->general_list.as<uint32_t>(), c->big_list.as<uint32_t>() + An, c->vals.as<double>(), V,
                c->arena.as<uint8_t>(), acap, c->rec_ref.as<uint32_t>(), &dctr->gen,
                (use_lookup && c->lut_mi.cx3.p && !getenv("RIN_NO_MI3_START")) ? c->lut_mi.cx3.as<MIComplex<MICapsSmall>>()
                                                                                : nullptr,
                c->lut_mi.lut3cx.as<uint32_t>());
            general_mi_big_kernel<W><<<sm * 4, GEN_THREADS, 0, s>>>(c->tets.as<uint4>(),
                c->act_tet.as<uint32_t>(), c->act_mask.as<uint32_t>(), c->act_cap, c->big_list.as<uint32_t>(),
                c->big_list.as<uint32_t>() + An, c->vals.as<double>(), V, c->arena.as<uint8_t>(), acap,
                c->rec_ref.as<uint32_t>(), &dctr->gen);
            CK(cudaGetLastError());
            // counts + offsets ride behind the general kernels (a record that did not fit reads as the empty record)
            EVREC(c->ev[ST_SCAN]);
            CK(cudaMemsetAsync(c->status.p, 0, (size_t)a_tiles * 16, s));
            if (attempt) CK(cudaMemsetAsync(&dctr->scan, 0, sizeof(ScanTotals), s));
            count_scan_kernel<W><<<a_tiles, 256, 0, s>>>(c->rec_ref.as<uint32_t>(), c->act_mask.as<uint32_t>(),
                c->act_cap, An, c->lut_mi.blob.as<uint8_t>(), c->arena.as<uint8_t>(), c->offs.as<uint4>(),
                c->status.as<unsigned long long>(), c->status.as<unsigned long long>() + a_tiles, &dctr->scan, dA);
            CK(cudaGetLastError());
            c->launches += 3;
            {
                // one copy of the whole counter block into pinned memory (copies into pageable memory serialise)
                Counters* hp = static_cast<Counters*>(c->h_pinned);
                CK(cudaMemcpyAsync(hp, dctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                h.gen = hp->gen;
                h.scan = hp->scan;
                n_gated = hp->n_tie_faces;
                if (nosync) {
                    h.filt = hp->filt;
                    h.n_zero = hp->n_zero;
                }
            }
            if (nosync && h.filt.n_active > An) { // more active tets than the learnt capacity: once more, sized exactly
                c->h_act_mi = 0;
                return run_mi_w<W>(c, flags);
            }
            if (h.gen.err)
                return fail(h.gen.err, "per-tet arrangement failed in tet " + std::to_string(h.gen.err_tet) +
                                           (h.gen.err == RIN_ERR_CAPACITY ? " (complex exceeds kernel capacity)"
                                                                         : " (degenerate input)"));
            if (!h.gen.arena_overflow) break;
            if (attempt > 2) return fail(RIN_ERR_STATE, "general kernel: arena overflow after regrow");
            CK(c->arena.ensure((size_t)h.gen.arena_top + h.gen.arena_top / 8 + 4096));
            GeneralCounters z{};
            z.n_general = h.gen.n_general;
            z.n_small = h.gen.n_small;
            z.n_big = h.gen.n_big;
            CK(cudaMemcpyAsync(&dctr->gen, &z, sizeof(GeneralCounters), cudaMemcpyHostToDevice, s));
        }
    }
    if (nosync) A = h.filt.n_active;
    rin_counts& n = c->counts;
    n = rin_counts{};
    n.num_pts = c->V;
    n.num_tets = c->t_count;
    n.num_funcs = F;
    n.num_degenerate_vertex = h.n_zero;
    n.num_intersecting_tet = A;
    n.num_k1 = h.filt.n_k1;
    n.num_k2 = h.filt.n_k2;
    n.num_kmore = h.filt.n_kmore;
    n.num_active_funcs = h.filt.n_funcs;
    n.num_general_tets = h.gen.n_general;
    c->h_act_mi = A + A / 8 + 1024;
    c->h_gen_mi = h.gen.n_general + h.gen.n_general / 8 + 64;
    c->act_cap = std::max<uint32_t>(c->act_cap, c->h_act_mi); // (the row stride only ever grows)

    if (!An) EVREC(c->ev[ST_SCAN]);
    const uint32_t NC = h.scan.n_cand, NFc = h.scan.n_faces, NFV = h.scan.n_fv;

    // ---- K5b: emit
    EVREC(c->ev[ST_EMIT]);
    CK(c->cand_key.ensure((size_t)std::max(NC, 1u) * 16));
    CK(c->cand_pay.ensure((size_t)std::max(NC, 1u) * 16));
    CK(c->face_hdr.ensure((size_t)std::max(NFc, 1u) * 16));
    CK(c->fv_ref.ensure((size_t)std::max(NFV, 1u) * 4));
    // tie tets exist: boundary faces carry the material set of their inside cell
    uint32_t* bf_mask = nullptr;
    if (n_gated) {
        CK(c->bf_mask.ensure((size_t)std::max(NFc, 1u) * W * 4));
        bf_mask = c->bf_mask.as<uint32_t>();
    }
    if (A) {
        emit_mi_kernel<W><<<grid_for(A, 256, sm, 4), 256, 0, s>>>(c->tets.as<uint4>(), c->act_tet.as<uint32_t>(),
            c->act_mask.as<uint32_t>(), c->act_cap, A, c->rec_ref.as<uint32_t>(), c->offs.as<uint4>(),
            c->lut_mi.blob.as<uint8_t>(), c->arena.as<uint8_t>(), c->cand_key.as<uint4>(),
            c->cand_pay.as<uint4>(), c->face_hdr.as<uint4>(), c->fv_ref.as<uint32_t>(), bf_mask,
            &dctr->n_bndry_faces);
        CK(cudaGetLastError());
        ++c->launches;
    }

    // ---- K6: dedup (hash-min + rank)
    EVREC(c->ev[ST_DEDUP]);
    uint32_t NV = 0;
    if (NC) {
        uint32_t tsize = 1024;
        while (tsize < 2 * NC) tsize <<= 1;
        CK(c->table.ensure((size_t)tsize * 4));
        CK(c->slot_of.ensure((size_t)NC * 4));
        CK(c->rep.ensure((size_t)NC * 4));
        CK(c->vid.ensure((size_t)NC * 4));
        CK(cudaMemsetAsync(c->table.p, 0xff, (size_t)tsize * 4, s));
        hash_insert_kernel<<<grid_for(NC, 256, sm, 8), 256, 0, s>>>(c->cand_key.as<uint4>(),
            c->cand_pay.as<uint4>(), NC, c->table.as<uint32_t>(), tsize - 1, c->slot_of.as<uint32_t>());
        // degenerate ties: reserved boundary-face slots exist -> match them across tets (only tie tets reserve
        // slots: without them nothing has to be read back here)
        unsigned n_slots = 0;
        if (n_gated) {
            CK(cudaMemcpyAsync(&n_slots, &dctr->n_bndry_faces, 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
        h.n_bndry_faces = n_slots;
        if (n_slots) {
            uint32_t t2 = 1024;
            while (t2 < 2 * NFc) t2 <<= 1;
            CK(c->ftable.ensure((size_t)t2 * 4));
            CK(c->bfkeys.ensure((size_t)NFc * 16));
            CK(c->frep.ensure((size_t)NFc * 4));
            CK(c->fpos.ensure((size_t)NFc * 16));
            CK(c->fdup.ensure((size_t)NFc * 8 + 16));
            CK(c->bids.ensure((size_t)NFc * 4)); // face slot -> table slot
            uint32_t* ndup = c->fdup.as<uint32_t>();
            CK(cudaMemsetAsync(c->ftable.p, 0xff, (size_t)t2 * 4, s));
            CK(cudaMemsetAsync(c->fdup.p, 0, (size_t)NFc * 8 + 16, s));
            CK(cudaMemsetAsync(&dctr->n_unique, 0, 4, s));
            const int g = grid_for(NFc, 256, sm, 8);
            mi_bface_keys_kernel<<<g, 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc, c->fv_ref.as<uint32_t>(),
                c->cand_key.as<uint4>(), c->cand_pay.as<uint4>(), c->slot_of.as<uint32_t>(), c->table.as<uint32_t>(),
                c->bfkeys.as<uint4>());
            bface_insert_kernel<<<g, 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc, c->bfkeys.as<uint4>(),
                c->ftable.as<uint32_t>(), t2 - 1, c->bids.as<uint32_t>());
            bface_reps_kernel<<<g, 256, 0, s>>>(c->ftable.as<uint32_t>(), c->bids.as<uint32_t>(), NFc,
                c->frep.as<uint32_t>(), ndup);
            CK(c->fpartner.ensure((size_t)NFc * 4));
            CK(cudaMemsetAsync(c->fpartner.p, 0xff, (size_t)NFc * 4, s));
            mi_bface_decide_kernel<<<g, 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc, c->frep.as<uint32_t>(), ndup,
                c->fv_ref.as<uint32_t>(), c->cand_pay.as<uint4>(), bf_mask, c->act_tet.as<uint32_t>(),
                c->act_mask.as<uint32_t>(), c->act_cap, A, W, &dctr->n_unique, c->fpartner.as<uint32_t>());
            // the activated corner candidates join the vertex table
            hash_insert_kernel<<<grid_for(NC, 256, sm, 8), 256, 0, s>>>(c->cand_key.as<uint4>(),
                c->cand_pay.as<uint4>(), NC, c->table.as<uint32_t>(), tsize - 1, c->slot_of.as<uint32_t>());
            CK(cudaGetLastError());
            c->launches += 5;
            unsigned bad = 0;
            CK(cudaMemcpyAsync(&bad, &dctr->n_unique, 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (bad) return fail(RIN_ERR_STATE, "material interface: a tet face is shared by more than two active tets");
        }
        const uint32_t r_tiles = (NC + 1023) / 1024;
        CK(c->status.ensure((size_t)r_tiles * 8 + 64));
        CK(cudaMemsetAsync(c->status.p, 0, (size_t)r_tiles * 8, s));
        rank_reps_kernel<<<r_tiles, 256, 0, s>>>(c->table.as<uint32_t>(), c->slot_of.as<uint32_t>(),
            c->cand_pay.as<uint4>(), NC,
            c->rep.as<uint32_t>(), c->vid.as<uint32_t>(), c->status.as<unsigned long long>(), &dctr->rank_tile,
            &dctr->n_unique);
        CK(cudaGetLastError());
        c->launches += 2; // hash_insert, rank_reps
    }

    // ---- K7: unique vertices + xyz (their number arrives with the final read-back: sized for the candidates)
    EVREC(c->ev[ST_VERTS]);
    CK(c->v_tet.ensure((size_t)std::max(NC, 1u) * 4));
    CK(c->v_local.ensure(std::max(NC, 1u)));
    CK(c->v_size.ensure(std::max(NC, 1u)));
    CK(c->v_simplex.ensure((size_t)std::max(NC, 1u) * 16));
    CK(c->v_funcs.ensure((size_t)std::max(NC, 1u) * 16));
    CK(c->v_xyz.ensure((size_t)std::max(NC, 1u) * 24));
    CK(c->v_key.ensure((size_t)std::max(NC, 1u) * 16));
    if (NC) {
        write_verts_mi_kernel<<<grid_for(NC, 256, sm, 8), 256, 0, s>>>(c->cand_key.as<uint4>(),
            c->cand_pay.as<uint4>(), c->rep.as<uint32_t>(), c->vid.as<uint32_t>(), NC, c->tets.as<uint4>(),
            c->vals.as<double>(), V, c->pts.as<double>(), c->v_tet.as<uint32_t>(), c->v_local.as<uint8_t>(),
            c->v_size.as<uint8_t>(), c->v_simplex.as<uint4>(), c->v_funcs.as<uint4>(), c->v_xyz.as<double>(),
            c->v_key.as<uint4>());
        CK(cudaGetLastError());
        ++c->launches;
    }

    // ---- faces
    EVREC(c->ev[ST_FACES]);
    CK(c->f_off.ensure((size_t)(NFc + 1) * 4));
    CK(c->f_verts.ensure((size_t)std::max(NFV, 1u) * 4));
    CK(c->f_toff.ensure((size_t)(NFc + 1) * 4));
    CK(c->f_tets.ensure((size_t)std::max(NFc, 1u) * 8));
    CK(c->f_funcs.ensure((size_t)std::max(NFc, 1u) * 8));
    uint32_t NF = NFc, NFVout = NFV, NFT = NFc;
    if (!h.n_bndry_faces) {
        if (NFV)
            remap_face_verts_kernel<<<grid_for(NFV, 256, sm, 8), 256, 0, s>>>(c->fv_ref.as<uint32_t>(), NFV,
                c->rep.as<uint32_t>(), c->vid.as<uint32_t>(), c->f_verts.as<uint32_t>());
        write_faces_mi_kernel<<<grid_for((uint64_t)NFc + 1, 256, sm, 8), 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc,
            NFV, c->f_off.as<uint32_t>(), c->f_toff.as<uint32_t>(), c->f_tets.as<uint32_t>(),
            c->f_funcs.as<uint32_t>());
        CK(cudaGetLastError());
        c->launches += NFV ? 2 : 1;
    } else {
        // drop the boundary-face slots that were not switched on
        CK(c->tmp_fverts.ensure((size_t)std::max(NFV, 1u) * 4));
        uint32_t* ndup = c->fdup.as<uint32_t>();
        uint32_t* cursor = ndup + NFc;
        uint32_t* totals = cursor + NFc;
        CK(cudaMemsetAsync(c->fdup.p, 0, (size_t)NFc * 8 + 16, s));
        remap_face_verts_kernel<<<grid_for(NFV, 256, sm, 8), 256, 0, s>>>(c->fv_ref.as<uint32_t>(), NFV,
            c->rep.as<uint32_t>(), c->vid.as<uint32_t>(), c->tmp_fverts.as<uint32_t>());
        mi_face_keep_kernel<<<grid_for(NFc, 256, sm, 8), 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc,
            c->frep.as<uint32_t>());
        bface_scan_kernel<<<1, 1024, 0, s>>>(c->face_hdr.as<uint4>(), NFc, c->frep.as<uint32_t>(), ndup,
            c->fpos.as<uint4>(), totals);
        bface_write_kernel<<<grid_for((uint64_t)NFc + 1, 256, sm, 8), 256, 0, s>>>(c->face_hdr.as<uint4>(), NFc,
            c->tmp_fverts.as<uint32_t>(), c->frep.as<uint32_t>(), c->fpos.as<uint4>(), cursor, totals,
            c->f_off.as<uint32_t>(), c->f_verts.as<uint32_t>(), c->f_toff.as<uint32_t>(),
            c->f_tets.as<uint32_t>(), c->f_funcs.as<uint32_t>(), 1);
        CK(cudaGetLastError());
        c->launches += 4;
        uint32_t ht[3];
        CK(cudaMemcpyAsync(ht, totals, 12, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        NF = ht[0];
        NFVout = ht[1];
        NFT = ht[2];
    }
    EVREC(c->ev[ST_COUNT]);
    {
        Counters* hp = static_cast<Counters*>(c->h_pinned);
        CK(cudaMemcpyAsync(hp, dctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (NC) NV = hp->n_unique;
        h.n_exact_classify = hp->n_exact_classify;
        h.gen.n_exact = hp->gen.n_exact;
    }
    if (c->stage_timing) {
        for (int i = 0; i < ST_COUNT; ++i) CK(cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]));
        CK(cudaEventElapsedTime(&c->kernel_ms[0], c->kev[0], c->kev[1]));
        CK(cudaEventElapsedTime(&c->kernel_ms[1], c->kev[2], c->kev[3]));
    } else {
        for (auto& x : c->stage_ms) x = 0;
        c->kernel_ms[0] = c->kernel_ms[1] = 0;
    }
    CK(cudaEventElapsedTime(&c->total_ms, c->ev[0], c->ev[ST_COUNT]));

    c->marked = c->finalized = false;
    c->ia_bndry_faces = h.n_bndry_faces != 0;
    c->maps_ready = false;
    c->n_local_verts = NV;
    c->n_own = NV;
    n.num_verts = NV;
    n.num_faces = NF;
    n.num_face_verts = NFVout;
    n.num_face_tets = NFT;
    n.num_exact_fallbacks = (uint64_t)h.gen.n_exact + h.n_exact_classify;
    return RIN_OK;
}

// ------------------------------------------------------------------------------------------------
// Lookup tables (1 and 2 functions), generated by THIS library's general kernel on witness tets:
// one witness per vertex-sign pattern (1 function) and per (sign pattern, crossing order) key
// (2 functions).  A key that no witness realises stays LUT_MISS and takes the general kernel.
// ------------------------------------------------------------------------------------------------
uint64_t splitmix(uint64_t& s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
double urand(uint64_t& s)
{
    return (splitmix(s) >> 11) * (1.0 / 9007199254740992.0);
}

int build_ia_tables(rin_ctx* c)
{
    cudaStream_t s = c->stream;
    const int sm = c->sm_count;
    Lut& L = c->lut_ia;
    // witnesses: tet w has private vertices 4w..4w+3; functions 0,1
    const uint32_t NW1 = 16, NW2 = 1u << 18, NWT = NW1 + NW2;
    const uint32_t Vw = 4 * NWT;
    std::vector<double> vals((size_t)2 * Vw, 0.0);
    uint64_t seed = 0x5eed1a2b3c4dULL;
    for (uint32_t w = 0; w < NW1; ++w)
        for (int cidx = 0; cidx < 4; ++cidx) {
            vals[4 * w + cidx] = (((w >> cidx) & 1) ? 1.0 : -1.0) * (0.5 + urand(seed));
            vals[(size_t)Vw + 4 * w + cidx] = 1.0; // second function inactive (all positive)
        }
    for (uint32_t w = NW1; w < NWT; ++w) {
        uint32_t outer = (uint32_t)(splitmix(seed) & 255);
        for (int cidx = 0; cidx < 4; ++cidx) {
            vals[4 * w + cidx] = (((outer >> cidx) & 1) ? 1.0 : -1.0) * (0.02 + urand(seed));
            vals[(size_t)Vw + 4 * w + cidx] = (((outer >> (4 + cidx)) & 1) ? 1.0 : -1.0) * (0.02 + urand(seed));
        }
    }
    std::vector<uint4> tets(NWT);
    for (uint32_t w = 0; w < NWT; ++w) tets[w] = make_uint4(4 * w, 4 * w + 1, 4 * w + 2, 4 * w + 3);
    // masks
    std::vector<uint2> vm(Vw);
    for (uint32_t v = 0; v < Vw; ++v) {
        uint32_t P = 0, N = 0;
        for (int f = 0; f < 2; ++f) {
            double x = vals[(size_t)f * Vw + v];
            P |= (x > 0 ? 1u : 0u) << f;
            N |= (x < 0 ? 1u : 0u) << f;
        }
        vm[v] = make_uint2(P, N);
    }
    DevBuf d_vals, d_tets, d_vm, d_act_tet, d_act_mask, d_ref, d_gl, d_keys, d_ctr, d_arena, d_l1, d_l2;
    auto cleanup = [&]() {
        for (DevBuf* b : {&d_vals, &d_tets, &d_vm, &d_act_tet, &d_act_mask, &d_ref, &d_gl, &d_keys, &d_ctr,
                 &d_arena, &d_l1, &d_l2})
            b->release();
    };
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            cleanup();                                                                             \
            return fail(RIN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
        }                                                                                          \
    } while (0)
    CKC(d_vals.ensure(vals.size() * 8));
    CKC(d_tets.ensure((size_t)NWT * 16));
    CKC(d_vm.ensure((size_t)Vw * 8));
    CKC(d_act_tet.ensure((size_t)NWT * 4));
    CKC(d_act_mask.ensure((size_t)NWT * 4));
    CKC(d_ref.ensure((size_t)NWT * 4));
    CKC(d_gl.ensure((size_t)NWT * 4));
    CKC(d_keys.ensure((size_t)NWT * 4));
    CKC(d_ctr.ensure(sizeof(Counters)));
    CKC(d_l1.ensure(32));
    CKC(d_l2.ensure(256 * 64 * 2));
    CKC(cudaMemcpyAsync(d_vals.p, vals.data(), vals.size() * 8, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_tets.p, tets.data(), (size_t)NWT * 16, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_vm.p, vm.data(), (size_t)Vw * 8, cudaMemcpyHostToDevice, s));
    // every witness is "active"; mask = functions crossing it
    std::vector<uint32_t> at(NWT), am(NWT);
    for (uint32_t w = 0; w < NWT; ++w) {
        at[w] = w;
        am[w] = (w < NW1) ? 1u : 3u;
    }
    CKC(cudaMemcpyAsync(d_act_tet.p, at.data(), (size_t)NWT * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_act_mask.p, am.data(), (size_t)NWT * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemsetAsync(d_ctr.p, 0, sizeof(Counters), s));
    CKC(cudaMemsetAsync(d_l1.p, 0xff, 32, s));
    CKC(cudaMemsetAsync(d_l2.p, 0xff, 256 * 64 * 2, s));
    Counters* dctr = d_ctr.as<Counters>();
    // pass 1: keys of all witnesses (tables are all-miss, so nothing is looked up)
    LutView lv{d_l1.as<uint16_t>(), d_l2.as<uint16_t>(), nullptr, 0};
    classify_ia_kernel<1><<<grid_for(NWT, 256, sm, 4), 256, 0, s>>>(d_tets.as<uint4>(), d_act_tet.as<uint32_t>(),
        d_act_mask.as<uint32_t>(), NWT, NWT, d_vm.as<uint2>(), d_vals.as<double>(), Vw, lv, 1, 0,
        d_ref.as<uint32_t>(), d_gl.as<uint32_t>(), d_gl.as<uint32_t>(), &dctr->gen, &dctr->n_exact_classify,
        d_keys.as<int>());
    CKC(cudaGetLastError());
    std::vector<int> keys(NWT);
    CKC(cudaMemcpyAsync(keys.data(), d_keys.p, (size_t)NWT * 4, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    // first witness of every key
    std::vector<uint32_t> chosen;
    std::vector<int> chosen_key;
    {
        std::vector<char> seen1(16, 0), seen2(256 * 64, 0);
        for (uint32_t w = 0; w < NWT; ++w) {
            int k = keys[w];
            if (k < 0) continue;
            char& sn = (w < NW1) ? seen1[k] : seen2[k];
            if (sn) continue;
            sn = 1;
            chosen.push_back(w);
            chosen_key.push_back(w < NW1 ? k : (k | (1 << 20)));
        }
    }
    // pass 2: general kernel on the chosen witnesses
    const uint32_t NCH = (uint32_t)chosen.size();
    CKC(cudaMemcpyAsync(d_gl.p, chosen.data(), (size_t)NCH * 4, cudaMemcpyHostToDevice, s));
    GeneralCounters g0{};
    g0.n_general = NCH;
    g0.n_big = NCH;
    g0.arena_top = 8;
    CKC(cudaMemcpyAsync(&dctr->gen, &g0, sizeof(g0), cudaMemcpyHostToDevice, s));
    CKC(d_arena.ensure((size_t)NCH * 256 + 4096));
    CKC(cudaMemsetAsync(d_arena.p, 0, 8, s));
    general_ia_big_kernel<1><<<sm * 4, GEN_THREADS, 0, s>>>(d_tets.as<uint4>(), d_act_tet.as<uint32_t>(),
        d_act_mask.as<uint32_t>(), NWT, d_gl.as<uint32_t>(), &dctr->gen.n_big, d_vals.as<double>(), Vw,
        d_arena.as<uint8_t>(), (uint32_t)d_arena.cap, d_ref.as<uint32_t>(), &dctr->gen);
    CKC(cudaGetLastError());
    GeneralCounters g1;
    CKC(cudaMemcpyAsync(&g1, &dctr->gen, sizeof(g1), cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    if (g1.err || g1.arena_overflow) {
        cleanup();
        return fail(RIN_ERR_STATE, "table generation failed");
    }
    std::vector<uint8_t> arena(g1.arena_top);
    std::vector<uint32_t> refs(NWT);
    CKC(cudaMemcpyAsync(arena.data(), d_arena.p, g1.arena_top, cudaMemcpyDeviceToHost, s));
    CKC(cudaMemcpyAsync(refs.data(), d_ref.p, (size_t)NWT * 4, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    // pack: blob = [empty record][records in (table, key) order]
    L.h_lut1.assign(16, LUT_MISS);
    L.h_lut2.assign(256 * 64, LUT_MISS);
    L.h_blob.assign(8, 0); // empty record: header + trailing word
    std::map<int, uint32_t> order; // key (with table tag) -> witness
    for (uint32_t i = 0; i < NCH; ++i) order[chosen_key[i]] = chosen[i];
    for (auto& kv : order) {
        const uint32_t w = kv.second;
        const uint8_t* r = arena.data() + (size_t)(refs[w] & ~REF_GENERAL) * 4;
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(r);
        const int nv = rw[0] & 255, nf = (rw[0] >> 8) & 255;
        uint32_t words = 1 + nv;
        for (int f = 0; f < nf; ++f) words += rec_face_words((rw[words] >> 24) & 127);
        words += 1; // trailing word: faces of the whole complex
        const uint32_t sz = 4 * words;
        const uint32_t off = (uint32_t)L.h_blob.size() / 4;
        if (off >= LUT_MISS) {
            cleanup();
            return fail(RIN_ERR_STATE, "table blob too large");
        }
        L.h_blob.insert(L.h_blob.end(), r, r + sz);
        while (L.h_blob.size() % 4) L.h_blob.push_back(0);
        if (kv.first & (1 << 20))
            L.h_lut2[kv.first & 0xfffff] = (uint16_t)off;
        else
            L.h_lut1[kv.first] = (uint16_t)off;
    }
    // complete 2-plane complexes, one per 2-plane key, in key order
    {
        std::vector<uint32_t> w2;
        std::vector<uint16_t> idx(256 * 64, LUT_MISS);
        for (auto& kv : order)
            if (kv.first & (1 << 20)) {
                idx[kv.first & 0xfffff] = (uint16_t)w2.size();
                w2.push_back(kv.second);
            }
        CKC(L.cx2.ensure(std::max<size_t>(w2.size(), 1) * sizeof(IAComplex<IACapsSmall>)));
        CKC(L.lut2cx.ensure(256 * 64 * 2));
        CKC(cudaMemcpyAsync(d_gl.p, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice, s));
        int* d_err = reinterpret_cast<int*>(&dctr->gen.err);
        CKC(cudaMemsetAsync(d_err, 0, 4, s));
        if (!w2.empty())
            dump_ia2_kernel<<<sm, GEN_THREADS, 0, s>>>(d_gl.as<uint32_t>(), (uint32_t)w2.size(), d_vals.as<double>(),
                Vw, L.cx2.as<IAComplex<IACapsSmall>>(), d_err);
        CKC(cudaGetLastError());
        int herr = 0;
        CKC(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, s));
        CKC(cudaMemcpyAsync(L.lut2cx.p, idx.data(), 256 * 64 * 2, cudaMemcpyHostToDevice, s));
        CKC(cudaStreamSynchronize(s));
        if (herr) {
            cleanup();
            return fail(RIN_ERR_STATE, "table generation failed (2-plane complexes)");
        }
    }
    L.blob_bytes = (uint32_t)L.h_blob.size();
    CKC(L.lut1.ensure(32));
    CKC(L.lut2.ensure(256 * 64 * 2));
    CKC(L.blob.ensure(L.blob_bytes));
    CKC(cudaMemcpyAsync(L.lut1.p, L.h_lut1.data(), 32, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(L.lut2.p, L.h_lut2.data(), 256 * 64 * 2, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(L.blob.p, L.h_blob.data(), L.blob_bytes, cudaMemcpyHostToDevice, s));
    CKC(cudaStreamSynchronize(s));
    cleanup();
    L.built = true;
    return RIN_OK;
#undef CKC
}

// 16-entry table for two materials (sign pattern of m0 - m1 at the corners), generated by this
// library's general MI kernel on witness tets.
int build_mi_tables(rin_ctx* c)
{
    cudaStream_t s = c->stream;
    const int sm = c->sm_count;
    Lut& L = c->lut_mi;
    const uint32_t NW = 16, Vw = 4 * NW;
    std::vector<double> vals((size_t)2 * Vw);
    for (uint32_t w = 0; w < NW; ++w)
        for (int i = 0; i < 4; ++i) {
            const double a = 0.25 * i;
            vals[4 * w + i] = a;
            vals[(size_t)Vw + 4 * w + i] = a + (((w >> i) & 1) ? -1.0 : 1.0) * (0.5 + 0.125 * i);
        }
    std::vector<uint4> tets(NW);
    std::vector<uint32_t> at(NW), am(NW, 3u), gl(NW);
    for (uint32_t w = 0; w < NW; ++w) {
        tets[w] = make_uint4(4 * w, 4 * w + 1, 4 * w + 2, 4 * w + 3);
        at[w] = gl[w] = w;
    }
    DevBuf d_vals, d_tets, d_at, d_am, d_ref, d_gl, d_ctr, d_arena;
    auto cleanup = [&]() {
        for (DevBuf* b : {&d_vals, &d_tets, &d_at, &d_am, &d_ref, &d_gl, &d_ctr, &d_arena}) b->release();
    };
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            cleanup();                                                                             \
            return fail(RIN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
        }                                                                                          \
    } while (0)
    CKC(d_vals.ensure(vals.size() * 8));
    CKC(d_tets.ensure(NW * 16));
    CKC(d_at.ensure(NW * 4));
    CKC(d_am.ensure(NW * 4));
    CKC(d_ref.ensure(NW * 4));
    CKC(d_gl.ensure(NW * 4));
    CKC(d_ctr.ensure(sizeof(Counters)));
    CKC(d_arena.ensure(1 << 16));
    CKC(cudaMemcpyAsync(d_vals.p, vals.data(), vals.size() * 8, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_tets.p, tets.data(), NW * 16, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_at.p, at.data(), NW * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_am.p, am.data(), NW * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(d_gl.p, gl.data(), NW * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemsetAsync(d_ctr.p, 0, sizeof(Counters), s));
    CKC(cudaMemsetAsync(d_arena.p, 0, 4, s));
    CKC(cudaMemsetAsync(d_ref.p, 0, NW * 4, s)); // the general kernel reads the tie-gate flag from here
    Counters* dctr = d_ctr.as<Counters>();
    GeneralCounters g0{};
    g0.n_general = g0.n_big = NW;
    g0.arena_top = 4;
    CKC(cudaMemcpyAsync(&dctr->gen, &g0, sizeof(g0), cudaMemcpyHostToDevice, s));
    general_mi_big_kernel<1><<<sm, GEN_THREADS, 0, s>>>(d_tets.as<uint4>(), d_at.as<uint32_t>(), d_am.as<uint32_t>(),
        NW, d_gl.as<uint32_t>(), d_gl.as<uint32_t>(), d_vals.as<double>(), Vw, d_arena.as<uint8_t>(), 1 << 16,
        d_ref.as<uint32_t>(), &dctr->gen);
    CKC(cudaGetLastError());
    GeneralCounters g1;
    std::vector<uint32_t> refs(NW);
    CKC(cudaMemcpyAsync(&g1, &dctr->gen, sizeof(g1), cudaMemcpyDeviceToHost, s));
    CKC(cudaMemcpyAsync(refs.data(), d_ref.p, NW * 4, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    if (g1.err || g1.arena_overflow) {
        cleanup();
        return fail(RIN_ERR_STATE, "MI table generation failed");
    }
    std::vector<uint8_t> arena(g1.arena_top);
    CKC(cudaMemcpyAsync(arena.data(), d_arena.p, g1.arena_top, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    L.h_lut1.assign(16, 0);
    L.h_blob.assign(4, 0);
    for (uint32_t w = 0; w < NW; ++w) {
        if ((refs[w] & REF_FLAGS) != REF_GENERAL || (size_t)(refs[w] & ~REF_FLAGS) * 4 + 4 > arena.size()) {
            cleanup();
            return fail(RIN_ERR_STATE, "MI table generation: witness without a plain record");
        }
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(arena.data() + (size_t)(refs[w] & ~REF_FLAGS) * 4);
        const int nv = rw[0] & 255, nf = (rw[0] >> 8) & 255;
        uint32_t words = 1 + 2 * nv;
        for (int f = 0; f < nf; ++f) words += 1 + rec_face_words((rw[words] >> 24) & 127);
        words += 1; // trailing word: faces of the whole complex
        L.h_lut1[w] = (uint16_t)(L.h_blob.size() / 4);
        const uint8_t* r = reinterpret_cast<const uint8_t*>(rw);
        L.h_blob.insert(L.h_blob.end(), r, r + 4 * words);
    }
    // ---- three materials (the "secondary" table): hashed witnesses, the first witness of every key goes through
    // this library's general kernel; a key no witness realises stays LUT3_MISS and takes the general kernel
    {
        const uint32_t NW3 = 1u << 19, Vw3 = 4 * NW3;
        DevBuf d_v3, d_k3, d_t3, d_a3, d_m3, d_r3, d_g3, d_ar3;
        auto cleanup3 = [&]() {
            for (DevBuf* b : {&d_v3, &d_k3, &d_t3, &d_a3, &d_m3, &d_r3, &d_g3, &d_ar3}) b->release();
        };
#define CK3(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            cleanup3();                                                                            \
            cleanup();                                                                             \
            return fail(RIN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
        }                                                                                          \
    } while (0)
        CK3(d_v3.ensure((size_t)3 * Vw3 * 8));
        CK3(d_k3.ensure((size_t)NW3 * 4));
        mi3_witness_kernel<<<grid_for(NW3, 256, sm), 256, 0, s>>>(NW3, Vw3, d_v3.as<double>(), d_k3.as<int>());
        CK3(cudaGetLastError());
        std::vector<int> keys(NW3);
        CK3(cudaMemcpyAsync(keys.data(), d_k3.p, (size_t)NW3 * 4, cudaMemcpyDeviceToHost, s));
        CK3(cudaStreamSynchronize(s));
        std::map<int, uint32_t> first; // key -> first witness
        for (uint32_t w = 0; w < NW3; ++w)
            if (keys[w] >= 0) first.emplace(keys[w], w);
        const uint32_t NCH = (uint32_t)first.size();
        std::vector<uint4> t3(NCH);
        std::vector<uint32_t> a3(NCH), m3(NCH, 7u), g3(NCH);
        {
            uint32_t i = 0;
            for (auto& kv : first) {
                const uint32_t w = kv.second;
                t3[i] = make_uint4(4 * w, 4 * w + 1, 4 * w + 2, 4 * w + 3);
                a3[i] = g3[i] = i;
                ++i;
            }
        }
        CK3(d_t3.ensure((size_t)NCH * 16));
        CK3(d_a3.ensure((size_t)NCH * 4));
        CK3(d_m3.ensure((size_t)NCH * 4));
        CK3(d_r3.ensure((size_t)NCH * 4));
        CK3(d_g3.ensure((size_t)NCH * 4));
        CK3(d_ar3.ensure((size_t)NCH * 512 + 4096));
        CK3(cudaMemcpyAsync(d_t3.p, t3.data(), (size_t)NCH * 16, cudaMemcpyHostToDevice, s));
        CK3(cudaMemcpyAsync(d_a3.p, a3.data(), (size_t)NCH * 4, cudaMemcpyHostToDevice, s));
        CK3(cudaMemcpyAsync(d_m3.p, m3.data(), (size_t)NCH * 4, cudaMemcpyHostToDevice, s));
        CK3(cudaMemcpyAsync(d_g3.p, g3.data(), (size_t)NCH * 4, cudaMemcpyHostToDevice, s));
        CK3(cudaMemsetAsync(d_r3.p, 0, (size_t)NCH * 4, s));
        CK3(cudaMemsetAsync(d_ar3.p, 0, 4, s));
        GeneralCounters g3c{};
        g3c.n_general = g3c.n_big = NCH;
        g3c.arena_top = 4;
        CK3(cudaMemcpyAsync(&dctr->gen, &g3c, sizeof(g3c), cudaMemcpyHostToDevice, s));
        general_mi_big_kernel<1><<<sm * 4, GEN_THREADS, 0, s>>>(d_t3.as<uint4>(), d_a3.as<uint32_t>(),
            d_m3.as<uint32_t>(), NCH, d_g3.as<uint32_t>(), d_g3.as<uint32_t>(), d_v3.as<double>(), Vw3,
            d_ar3.as<uint8_t>(), (uint32_t)d_ar3.cap, d_r3.as<uint32_t>(), &dctr->gen);
        CK3(cudaGetLastError());
        GeneralCounters g3r;
        std::vector<uint32_t> refs3(NCH);
        CK3(cudaMemcpyAsync(&g3r, &dctr->gen, sizeof(g3r), cudaMemcpyDeviceToHost, s));
        CK3(cudaMemcpyAsync(refs3.data(), d_r3.p, (size_t)NCH * 4, cudaMemcpyDeviceToHost, s));
        CK3(cudaStreamSynchronize(s));
        if (g3r.err || g3r.arena_overflow) {
            cleanup3();
            cleanup();
            return fail(RIN_ERR_STATE, "MI 3-material table generation failed");
        }
        std::vector<uint8_t> ar3(g3r.arena_top);
        CK3(cudaMemcpyAsync(ar3.data(), d_ar3.p, g3r.arena_top, cudaMemcpyDeviceToHost, s));
        CK3(cudaStreamSynchronize(s));
        std::vector<uint32_t> lut3(MI3_KEYS, LUT3_MISS);
        uint32_t i = 0;
        for (auto& kv : first) {
            const uint32_t ref = refs3[i++];
            if ((ref & REF_FLAGS) != REF_GENERAL || (size_t)(ref & ~REF_FLAGS) * 4 + 4 > ar3.size()) continue;
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(ar3.data() + (size_t)(ref & ~REF_FLAGS) * 4);
            const int nv = rw[0] & 255, nf = (rw[0] >> 8) & 255;
            uint32_t words = 1 + 2 * nv;
            for (int f = 0; f < nf; ++f) words += 1 + rec_face_words((rw[words] >> 24) & 127);
            words += 1; // trailing word: faces of the whole complex
            lut3[kv.first] = (uint32_t)(L.h_blob.size() / 4);
            const uint8_t* r = reinterpret_cast<const uint8_t*>(rw);
            L.h_blob.insert(L.h_blob.end(), r, r + 4 * words);
        }
        CK3(L.lut2.ensure((size_t)MI3_KEYS * 4));
        CK3(cudaMemcpyAsync(L.lut2.p, lut3.data(), (size_t)MI3_KEYS * 4, cudaMemcpyHostToDevice, s));
        CK3(cudaStreamSynchronize(s));
        L.n_keys3 = NCH;
        // complete complexes of the same witnesses: tets with >= 4 materials start from them (general_mi_small_kernel)
        if (NCH) {
            std::vector<uint32_t> wit(NCH), lut3cx(MI3_KEYS, 0xffffffffu);
            uint32_t j = 0;
            for (auto& kv : first) {
                wit[j] = kv.second;
                if (lut3[kv.first] != LUT3_MISS) lut3cx[kv.first] = j;
                ++j;
            }
            CK3(d_g3.ensure((size_t)NCH * 4));
            CK3(cudaMemcpyAsync(d_g3.p, wit.data(), (size_t)NCH * 4, cudaMemcpyHostToDevice, s));
            CK3(L.cx3.ensure((size_t)NCH * sizeof(MIComplex<MICapsSmall>)));
            CK3(L.lut3cx.ensure((size_t)MI3_KEYS * 4));
            int* d_err = reinterpret_cast<int*>(&dctr->gen.err);
            CK3(cudaMemsetAsync(d_err, 0, 4, s));
            dump_mi3_kernel<<<sm * 4, GEN_THREADS, 0, s>>>(d_g3.as<uint32_t>(), NCH, d_v3.as<double>(), Vw3,
                L.cx3.as<MIComplex<MICapsSmall>>(), d_err);
            CK3(cudaGetLastError());
            int herr = 0;
            CK3(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, s));
            CK3(cudaMemcpyAsync(L.lut3cx.p, lut3cx.data(), (size_t)MI3_KEYS * 4, cudaMemcpyHostToDevice, s));
            CK3(cudaStreamSynchronize(s));
            if (herr) { // never seen; without the start complexes the general kernel inserts every material
                L.cx3.release();
                L.lut3cx.release();
            }
        }
        cleanup3();
#undef CK3
    }
    L.blob_bytes = (uint32_t)L.h_blob.size();
    CKC(L.lut1.ensure(32));
    CKC(L.blob.ensure(L.blob_bytes));
    CKC(cudaMemcpyAsync(L.lut1.p, L.h_lut1.data(), 32, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(L.blob.p, L.h_blob.data(), L.blob_bytes, cudaMemcpyHostToDevice, s));
    CKC(cudaStreamSynchronize(s));
    cleanup();
    L.built = true;
    return RIN_OK;
#undef CKC
}

} // namespace
