/* rin_b200 — C-ABI of the B200-native per-tetrahedron engine.
 *
 * This is the drop-in boundary for the hot path of Robust-Implicit-Surface-Networks:
 * function evaluation -> vertex signs -> per-tet active-function filter -> per-tet
 * arrangement -> mesh extraction (+ xyz).  Plain pointers and sizes only; every call
 * returns 0 on success and a negative code on failure (text via rin_last_error()).
 * No exceptions cross this boundary and there is NO CPU fallback: every entry point
 * that computes requires a CUDA device (RIN_ERR_NO_DEVICE otherwise).
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   rin_eval_functions*        load_functions()            app/implicit_arrangement.cpp:57-64
 *                                                          (un-vendored implicit_functions lib)
 *   rin_generate_grid          generate_tet_mesh()         src/io.cpp:95-152
 *   rin_run(IA)                implicit_arrangement() stages "func signs", "filter",
 *                              "simp_arr*", "extract mesh", "compute xyz"
 *                                                          src/implicit_arrangement.cpp:53-402
 *                              + compute_arrangement()     (un-vendored simplicial_arrangement)
 *                              + extract_iso_mesh()        src/extract_mesh.cpp:10-265
 *                              + compute_iso_vert_xyz()    src/extract_mesh.cpp:1446-1538
 *   rin_run(MI)                material_interface() stages "highest func", "filter", "MI*",
 *                              "extract mesh", "compute xyz" src/material_interface.cpp:53-447
 *                              + extract_MI_mesh()         src/extract_mesh.cpp:569-986
 *                              + compute_MI_vert_xyz()     src/extract_mesh.cpp:1541-1637
 *   rin_tet_maps               extract_iso_mesh() / extract_MI_mesh(), cell-grouping overloads:
 *                              global_vId_of_tet_vert, iso_fId_of_tet_face / MI_fId_of_tet_face
 *                                                          src/extract_mesh.cpp:268-566, :988-1443
 *   rin_get_complexes          cut_results[cut_result_index[tet]] as consumed by
 *                              src/pair_faces.cpp:138-238, src/topo_ray_shooting.cpp:56-57
 */
#ifndef RIN_B200_H
#define RIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rin_ctx rin_ctx;

enum {
    RIN_OK = 0,
    RIN_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product never falls back to CPU */
    RIN_ERR_CUDA = -2,
    RIN_ERR_ARG = -3,
    RIN_ERR_CAPACITY = -4,    /* a per-tet complex exceeded the kernel's fixed capacity */
    RIN_ERR_ARRANGEMENT = -5, /* per-tet computation failed (reference: std::runtime_error) */
    RIN_ERR_STATE = -6
};

/* function kinds (JSON "type" of the reference's function files) */
enum { RIN_FN_PLANE = 0, RIN_FN_SPHERE = 1, RIN_FN_CYLINDER = 2, RIN_FN_ZERO = 3, RIN_FN_TORUS = 4 };

/* parameters, by kind:
 *  PLANE    p[0..2] point, p[3..5] normal                      value  n.(x-point)
 *  SPHERE   p[0..2] center, p[3] radius, p[4] squared(0/1)     value  r-|x-c|  or  r^2-|x-c|^2
 *  CYLINDER p[0..2] axis_point, p[3..5] axis_vector (unit), p[6] radius    value r-dist(x,axis)
 *  TORUS    p[0..2] center, p[3..5] axis_vector (unit), p[6] major, p[7] minor
 *  ZERO     -                                                  value  0
 *  flip != 0 negates the value ("is_flipped").
 */
typedef struct rin_func_desc {
    int32_t type;
    int32_t flip;
    double p[10];
} rin_func_desc;

enum { RIN_MODE_IA = 0, RIN_MODE_MI = 1 };
#define RIN_MAX_FUNCS 128 /* functions / materials per run (four 32-bit mask words per tet) */
enum {
    RIN_FLAG_USE_LOOKUP = 1,           /* use_lookup            (src/implicit_arrangement.cpp:15) */
    RIN_FLAG_USE_SECONDARY_LOOKUP = 2, /* use_secondary_lookup  (:16, :38-40)                     */
    RIN_FLAG_NEGATE = 4                /* csg(): funcVals *= -1 (src/csg.cpp:37)                  */
};

/* stats.json fields produced by the hot path (src/implicit_arrangement.cpp:46-393,
 * src/material_interface.cpp:45-438) */
typedef struct rin_counts {
    uint64_t num_pts, num_tets, num_funcs;
    uint64_t num_degenerate_vertex; /* IA: (vertex,function) pairs with value == 0; MI: #vertices with tied maxima */
    uint64_t num_intersecting_tet;
    uint64_t num_k1, num_k2, num_kmore; /* IA: 1 / 2 / >=3 functions; MI: 2 / 3 / >=4 materials */
    uint64_t num_verts, num_faces;      /* iso (MI) vertices / faces after deduplication */
    uint64_t num_face_verts;            /* total length of the face vertex lists */
    uint64_t num_face_tets;             /* total (tet, local face) pairs */
    uint64_t num_general_tets;          /* tets that took the general (non-table) kernel */
    uint64_t num_exact_fallbacks;       /* predicate evaluations that needed exact arithmetic */
    uint64_t num_active_funcs;          /* length of func_in_tet */
} rin_counts;

/* host-side views filled by rin_download_mesh (caller allocates; sizes from rin_counts).
 * Any pointer may be NULL to skip that array.  Mesh_None is written as UINT64_MAX / UINT32_MAX. */
typedef struct rin_mesh_out {
    /* vertices (IsoVert / MI_Vert, src/mesh.h:28-57) */
    uint32_t* vert_tet;          /* [num_verts]     tet_index */
    uint8_t*  vert_local;        /* [num_verts]     tet_vert_index */
    uint8_t*  vert_simplex_size; /* [num_verts]     1..4 */
    uint32_t* vert_simplex;      /* [num_verts*4]   simplex_vert_indices (unused slots UINT32_MAX) */
    uint32_t* vert_funcs;        /* [num_verts*4]   func_indices[3] (IA) / material_indices[4] (MI) */
    double*   vert_xyz;          /* [num_verts*3] */
    /* faces (PolygonFace, src/mesh.h:15-25) */
    uint32_t* face_offsets;      /* [num_faces+1]   into face_verts */
    uint32_t* face_verts;        /* [num_face_verts] */
    uint32_t* face_tet_offsets;  /* [num_faces+1]   into face_tets */
    uint32_t* face_tets;         /* [num_face_tets*2] (tet, local face id) pairs */
    uint32_t* face_funcs;        /* [num_faces*2]   func_index.first / .second */
} rin_mesh_out;

const char* rin_last_error(void);
int rin_device_count(void);

int rin_create(int device, rin_ctx** out);
void rin_destroy(rin_ctx* ctx);

/* ---- inputs ---------------------------------------------------------------------------- */
/* explicit mesh, host pointers (pts: V*3 doubles; tets: T*4 indices, 64-bit as in the
 * reference's std::array<size_t,4>, or 32-bit) */
int rin_set_mesh_host(rin_ctx*, const double* pts, uint64_t n_pts, const void* tets,
                      uint64_t n_tets, int index_bytes /* 4 or 8 */);
/* one rank's slice of a mesh sharded by contiguous tet ranges: pts_slice = points [v_first, v_first + v_count)
 * (the vertices the tet range references), tets_slice = tets [t_first, t_first + t_count) with GLOBAL vertex
 * indices; only the slice crosses PCIe, the tet range of the run is set to the slice */
int rin_set_mesh_host_range(rin_ctx*, uint64_t n_pts, uint64_t n_tets, const double* pts_slice, uint64_t v_first,
                            uint64_t v_count, const void* tets_slice, uint64_t t_first, uint64_t t_count,
                            int index_bytes);
/* rows [v_first, v_first + v_count) of the row-major V x F value matrix */
int rin_set_values_host_range(rin_ctx*, const double* vals_slice, uint64_t v_first, uint64_t v_count,
                              uint32_t n_funcs);
/* generate_tet_mesh() on the device (same vertex and tet order as src/io.cpp:95-152) */
int rin_generate_grid(rin_ctx*, uint32_t resolution, const double bbox_min[3], const double bbox_max[3]);
/* restrict the run to tets [first, first+count) (slab sharding).  count == RIN_TET_RANGE_ALL: up to the last
 * tet; count == 0: an empty range (rin_run then returns an empty result); out-of-range -> RIN_ERR_ARG */
#define RIN_TET_RANGE_ALL UINT64_MAX
int rin_set_tet_range(rin_ctx*, uint64_t first, uint64_t count);
/* Ghost tets of a sharded run on DEGENERATE inputs (a function vanishing at grid vertices, materials tying there):
 * an iso-face can then lie on a tet face of the slab plane, where the reference pairs the two incident tets
 * (src/extract_mesh.cpp:240-253) or matches their materials (:833-981).  The run processes
 * [first - below, first + count + above) but keeps only what the rank's OWN tets create: vertices first created by a
 * ghost below are foreign (the lower rank owns them), everything created by a ghost above is dropped, faces keep
 * their complete tet lists.  `below` / `above` must cover every tet of the neighbouring ranks that touches a shared
 * vertex: one layer of cubes (5 R^2 tets) for x-slabs of a generated grid.  Call after rin_set_tet_range (which
 * resets the ghosts to 0); the values are clamped to the mesh.  rin_exchange_nccl enables one cube layer by itself
 * and repeats the run when any rank reports degenerate vertices on a generated grid. */
int rin_set_ghost_tets(rin_ctx*, uint64_t below, uint64_t above);

/* function values: either parametric (evaluated on the device) ... */
int rin_set_functions(rin_ctx*, const rin_func_desc* funcs, uint32_t n_funcs);
/* ... or given by the caller as the reference's row-major V x F matrix (host pointer) */
int rin_set_values_host(rin_ctx*, const double* vals_rowmajor, uint64_t n_pts, uint32_t n_funcs);

/* ---- run -------------------------------------------------------------------------------- */
int rin_run(rin_ctx*, int mode, uint32_t flags);
int rin_get_counts(const rin_ctx*, rin_counts* out);
int rin_download_mesh(rin_ctx*, rin_mesh_out* out);
/* func_in_tet / start_index_of_tet (src/implicit_arrangement.cpp:84-85) for the tet range:
 * start has count+1 entries; either pointer may be NULL */
int rin_download_active(rin_ctx*, uint32_t* func_in_tet, uint64_t* start_index_of_tet);
/* ids of the active tets of the last run, ascending (num_intersecting_tet entries): what a load balancer needs */
int rin_download_active_tets(rin_ctx*, uint32_t* tet_ids);
/* row-major V x F function values as evaluated on the device */
int rin_download_values(rin_ctx*, double* vals_rowmajor);
int rin_download_grid(rin_ctx*, double* pts, uint32_t* tets);
/* per-stage device times of the last run in milliseconds (CUDA events); names via rin_stage_name */
int rin_get_stage_times(const rin_ctx*, float* ms, int capacity);
/* per-stage CUDA events are recorded while this is on (default); off: only the whole pass is timed
 * (a dozen event records per pass cost more host time than some of the kernels take) */
int rin_set_stage_timing(rin_ctx*, int on);
/* kernels launched by the last rin_run (counted at the launch sites) */
int rin_get_launch_count(const rin_ctx*);
const char* rin_stage_name(int i);
/* CUDA-event durations of the last run: the eval and filter kernels alone, and the whole pass */
int rin_get_kernel_times(const rin_ctx*, float* eval_ms, float* filter_ms, float* total_ms);
int rin_num_stages(void);

/* ---- per-tet complexes on demand (host topology stages need O(#chains+#components) tets) --- */
/* Serialises the full complexes of the given tets.  Layout per tet (uint32 words):
 *  IA: nv, nf, nc, n_unique(0 if all planes unique), then nv*3 plane ids, then per face
 *      [supporting_plane, positive_cell, negative_cell, n, v_0..v_{n-1}], per cell [n, f_0..],
 *      then if n_unique: np, np group ids, np orientation flags.
 *  MI: nv, nf, nc, n_unique, nv*4 material ids, per face [pos_label, neg_label, n, v..],
 *      per cell [material_label, n, f..], then if n_unique: nm, nm group ids.
 * offsets[i]..offsets[i+1] delimit tet i in `words`.  Call with words==NULL to get the size. */
int rin_get_complexes(rin_ctx*, int mode, uint32_t flags, const uint64_t* tet_ids, uint64_t n,
                      uint64_t* offsets, uint32_t* words, uint64_t* n_words);

/* ---- cell-grouping maps: the second extract_iso_mesh / extract_MI_mesh overloads
 *      (src/extract_mesh.cpp:268-566, :988-1443) ----
 * For the last run (either mode), per ACTIVE tet a (in tet order; active_tets[a] is its tet id):
 *   vert_ids[vert_offsets[a] + j]  = global_vId_of_tet_vert of local vertex j of the tet's complex: the
 *                                    iso-vertex id, or -(grid vertex id)-1 for a tet corner (:402,518)
 *   face_ids[face_offsets[a] + j]  = iso_fId_of_tet_face of local face j, UINT32_MAX when the face is not on
 *                                    an iso-surface (:529-556); MI: MI_fId_of_tet_face, a boundary face that is a
 *                                    material interface between two tie tets has the same id in both (:1391-1392)
 * (The reference's T+1-entry start arrays follow by giving inactive tets empty ranges, :329-330.)
 * rin_tet_maps computes them on the device and returns the sizes; rin_download_tet_maps copies them
 * (offset arrays have n_active + 1 entries; any pointer may be NULL).  Single-process runs only. */
int rin_tet_maps(rin_ctx*, uint64_t* n_active, uint64_t* n_vert_entries, uint64_t* n_face_entries);
int rin_download_tet_maps(rin_ctx*, uint32_t* active_tets, uint32_t* vert_offsets, int64_t* vert_ids,
                          uint32_t* face_offsets, uint32_t* face_ids);

/* ---- N1: edges of the extracted mesh (compute_mesh_edges, src/mesh_connectivity.cpp:10-56) ----
 * Edge ids follow the first occurrence scanning (face, position in face).  edges_of_face[p] is the edge between
 * entries p and p + 1 (cyclically) of the face-vertex array; edge_faces lists, per edge and ascending, the
 * (face, position in face) pairs = Edge::face_edge_indices (src/mesh.h:62-71). */
typedef struct rin_edges_out {
    uint32_t* edge_verts;        /* [num_edges*2]      v1 <= v2 */
    uint32_t* edges_of_face;     /* [num_face_verts]   same offsets as face_verts */
    uint32_t* edge_face_offsets; /* [num_edges+1] */
    uint32_t* edge_faces;        /* [num_face_verts*2] (face, position) pairs */
} rin_edges_out;
int rin_mesh_edges(rin_ctx*, uint64_t* n_edges);
int rin_download_edges(rin_ctx*, rin_edges_out* out);

/* robust_test (-R) of the reference (src/implicit_arrangement.cpp:137-243): every active tet of the
 * last run is computed with its functions in forward and in reversed order.
 * out = {type 1 (inconsistent counts), type 2 (forward run failed), type 3 (reversed run failed), tets tested} */
int rin_robust_test(rin_ctx*, int mode, uint32_t out[4]);

/* ---- one-shot host entry (what the C++ drop-in drivers call) ---------------------------- */
int rin_run_host(rin_ctx*, int mode, uint32_t flags, const double* pts, uint64_t n_pts,
                 const void* tets, uint64_t n_tets, int index_bytes,
                 const double* vals_rowmajor, uint32_t n_funcs, rin_counts* counts);

/* ---- slab-boundary exchange for multi-GPU runs (one process per GPU) ----------------------
 * Every rank runs rin_run on its own contiguous tet range (rin_set_tet_range).  Vertices whose
 * minimal simplex lies in the vertex range shared with another rank are found by both; the rank
 * with the LOWER tet range owns them (it is the first occurrence in tet order).  Protocol, with
 * the collectives done by the caller (NCCL all-gather of a few KB):
 *   1. rin_boundary_export(own_only = 0, [lo, hi]) on every rank -> keys of shared candidates
 *   2. all-gather the keys; rin_mark_foreign(keys of all LOWER ranks) -> n_own
 *   3. all-gather n_own -> this rank's global vertex offset
 *   4. rin_boundary_export(own_only = 1) -> (key, own index) of the owned shared vertices;
 *      all-gather (key, offset + own index)
 *   5. rin_finalize_sharded(offset, foreign keys, foreign global ids): vertex arrays keep the
 *      owned vertices (first-occurrence order), face vertex lists are rewritten to global ids.
 * A key is four 32-bit words: sorted simplex vertex ids (unused 0xffffffff) and the packed
 * function / material ids. */
int rin_boundary_export(rin_ctx*, int own_only, uint32_t v_lo, uint32_t v_hi, uint32_t* keys, uint32_t* ids,
                        uint64_t capacity, uint64_t* n);
int rin_mark_foreign(rin_ctx*, const uint32_t* keys, uint64_t n_keys, uint64_t* n_own);
int rin_finalize_sharded(rin_ctx*, uint64_t vert_offset, const uint32_t* keys, const uint32_t* global_ids,
                         uint64_t n_keys);
/* The same protocol entirely on the device: stream-ordered kernels around two ncclAllGather calls
 * over NVLink / NVSwitch (NCCL is loaded at run time; env RIN_NCCL_LIB overrides the library name).
 * rank 0 creates the id, the caller distributes its 128 bytes, every rank calls rin_nccl_init. */
int rin_nccl_unique_id(uint8_t id[128]);
int rin_nccl_init(rin_ctx*, const uint8_t id[128], int rank, int world);
int rin_exchange_nccl(rin_ctx*, uint64_t* vert_offset, uint64_t* n_verts_total, uint64_t* face_offset,
                      uint64_t* n_faces_total);
/* rin_run + rin_exchange_nccl as one call with one host synchronisation (the exchange is enqueued behind the run's
 * kernels and reads its counts from device memory); falls back to the two calls when the fused path does not apply
 * (first pass over new inputs, material interface, degenerate inputs).  Collective: every rank calls it. */
int rin_run_exchange(rin_ctx*, int mode, uint32_t flags, uint64_t* vert_offset, uint64_t* n_verts_total,
                     uint64_t* face_offset, uint64_t* n_faces_total);
/* offsets of this rank's slice in the merged mesh after rin_exchange_nccl:
 * out = {vertex offset, vertices total, face offset, faces total, face-vertex offset, face-vertex total,
 *        face-tet-pair offset, face-tet-pair total}
 * After the exchange rin_download_mesh returns face_verts as global vertex ids and face_offsets /
 * face_tet_offsets rebased into the merged arrays: every rank can write its slice straight into shared arrays
 * of the merged sizes. */
int rin_get_exchange_offsets(const rin_ctx*, uint64_t out[8]);
/* device time (ms) of the exchange chain of the last fused rin_run_exchange (vertex kernel done -> owned vertices
 * compacted, including the wait for the neighbours); 0 when the pass was not fused */
int rin_get_exchange_time(const rin_ctx*, float* ms);
/* with stage timing on: device times (ms) of the kernels of the peer exchange chain {send keys, receive + insert, mark +
 * scan + publish record, offsets + global ids + owned vertices, 0, ...}, each from the end of the previous one */
int rin_get_exchange_parts(const rin_ctx*, float ms[10]);
/* vertex id range referenced by the current tet range */
int rin_get_vertex_range(const rin_ctx*, uint32_t* v_lo, uint32_t* v_hi);

#ifdef __cplusplus
}
#endif
#endif /* RIN_B200_H */
