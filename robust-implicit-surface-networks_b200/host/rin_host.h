// C++ host layer above the C-ABI (include/rin_b200.h): the hot stages of the reference's drivers
// with the reference's own container types, so that a maintainer can replace
// src/implicit_arrangement.cpp:53-402 / src/material_interface.cpp:53-447 by one call and keep the
// host topology stages (:404-647 / :449-695) unchanged.  No CPU fallback: every function returns
// false with a message when the CUDA library reports an error.
#pragma once
#include "mesh_types.h"

#include <cstdint>
#include <string>
#include <vector>

namespace rin_host {

// Per-tet complex as the reference's host stages read it (fields of simplicial_arrangement::
// Arrangement<3> / MaterialInterface<3>); None == SIZE_MAX.
struct TetComplex
{
    std::vector<std::array<size_t, 4>> vertices; // IA uses the first three entries
    struct Face
    {
        std::vector<size_t> vertices;
        size_t a = Mesh_None, b = Mesh_None, c = Mesh_None; // IA: supporting plane, positive cell, negative cell
                                                            // MI: positive label, negative label, -
    };
    std::vector<Face> faces;
    struct Cell
    {
        std::vector<size_t> faces;
        size_t material_label = Mesh_None;
    };
    std::vector<Cell> cells;
    std::vector<size_t> unique_indices;       // plane / material -> group (empty when all distinct)
    std::vector<bool> unique_orientations;    // IA only
};

struct HotPathOutput
{
    std::vector<size_t> func_in_tet, start_index_of_tet; // CRS of active functions / materials
    size_t num_degenerate_vertex = 0;                    // IA: zero (vertex, function) pairs
    size_t num_intersecting_tet = 0, num_k1 = 0, num_k2 = 0, num_kmore = 0;
    std::string error;
};

// funcVals: the reference's row-major V x F matrix (Eigen data()), rows == pts.size().
// Appends "func signs" .. "compute xyz" to timing_labels/timings and the stats of
// src/implicit_arrangement.cpp:46-393 to stats_labels/stats, exactly in the reference's order.
bool implicit_arrangement_hot(bool use_lookup, bool use_secondary_lookup,
    const std::vector<std::array<double, 3>>& pts, const std::vector<std::array<size_t, 4>>& tets,
    const double* funcVals, size_t n_func, bool negate, std::vector<std::array<double, 3>>& iso_pts,
    std::vector<PolygonFace>& iso_faces, std::vector<IsoVert>& iso_verts, HotPathOutput& out,
    std::vector<std::string>& timing_labels, std::vector<double>& timings, std::vector<std::string>& stats_labels,
    std::vector<size_t>& stats);

bool material_interface_hot(bool use_lookup, bool use_secondary_lookup,
    const std::vector<std::array<double, 3>>& pts, const std::vector<std::array<size_t, 4>>& tets,
    const double* funcVals, size_t n_func, std::vector<std::array<double, 3>>& MI_pts,
    std::vector<PolygonFace>& MI_faces, std::vector<MI_Vert>& MI_verts, HotPathOutput& out,
    std::vector<std::string>& timing_labels, std::vector<double>& timings, std::vector<std::string>& stats_labels,
    std::vector<size_t>& stats);

// robust_test (-R): forward / reversed insertion order on every active tet of the LAST hot-path call.
// Prints the reference's verdict line ("success." / "type N failure ...") and returns its value
// (src/implicit_arrangement.cpp:329-343).
bool robust_test_verdict(int mode, std::string& error);

// Complexes of the given tets of the LAST hot-path call (mode 0 = IA, 1 = MI); inactive tets give
// an empty complex.  This is the lazy replacement of cut_results[cut_result_index[tet]].
bool fetch_complexes(int mode, const std::vector<size_t>& tet_ids, std::vector<TetComplex>& out, std::string& error);

// number of per-tet complexes fetched since the last hot-path call (what "lazy" amounts to)
size_t complexes_fetched();

// Cell-grouping maps of the LAST implicit-arrangement call, in the shape the reference's second
// extract_iso_mesh overload returns them (src/extract_mesh.cpp:268-566) and build_simplicial_cell_adjacency
// consumes them (src/cell_connectivity.h:13-26): start arrays have n_tets + 1 entries.
bool fetch_tet_maps(size_t n_tets, std::vector<long long>& global_vId_of_tet_vert,
    std::vector<size_t>& global_vId_start_index_of_tet, std::vector<size_t>& iso_fId_of_tet_face,
    std::vector<size_t>& iso_fId_start_index_of_tet, std::string& error);

// ---- N1: the stages right behind the hot path (src/implicit_arrangement.cpp:406-465) -----------------------------
// Edges of the LAST hot-path call's mesh, computed on the device (rin_mesh_edges): the same ids, edges_of_face and
// face_edge_indices as compute_mesh_edges (src/mesh_connectivity.cpp:10-56).
bool mesh_edges(std::vector<std::vector<size_t>>& edges_of_face, std::vector<Edge>& mesh_edges, std::string& error);

// Faces grouped into patches across manifold edges (exactly two incident faces), in the reference's discovery
// order: patches by lowest face id, faces of a patch in breadth-first order, neighbours in the order of the
// face's edges (compute_patches, src/mesh_connectivity.cpp:58-94).  Label of a patch = label of its first face.
void mesh_patches(const std::vector<std::vector<size_t>>& edges_of_face, const std::vector<Edge>& mesh_edges,
    const std::vector<PolygonFace>& faces, std::vector<std::vector<size_t>>& patches,
    std::vector<size_t>& patch_function_label);
void mesh_patches(const std::vector<std::vector<size_t>>& edges_of_face, const std::vector<Edge>& mesh_edges,
    const std::vector<PolygonFace>& faces, std::vector<std::vector<size_t>>& patches,
    std::vector<std::pair<size_t, size_t>>& patch_function_label);

// Non-manifold edges (more than two incident faces) of every vertex, then chains: non-manifold edges linked
// through vertices with exactly two of them, breadth-first from the lowest edge id, end v1 before v2
// (src/implicit_arrangement.cpp:446-462, compute_chains src/mesh_connectivity.cpp:196-241).
void mesh_chains(size_t n_verts, const std::vector<Edge>& mesh_edges,
    std::vector<std::vector<size_t>>& non_manifold_edges_of_vert, std::vector<std::vector<size_t>>& chains);

} // namespace rin_host
