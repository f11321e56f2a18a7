// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of the public surface of the un-vendored third-party library
// qnzhou/simplicial_arrangement (GIT_TAG main, un-pinned;
// /root/reference/cmake/simplicial_arrangement.cmake:5-10) that the reference's in-tree
// sources dereference.  Field names/types are exactly those used at
// /root/reference/src/extract_mesh.cpp:54-91,630-651, src/pair_faces.cpp:143-237,
// src/topo_ray_shooting.cpp:56-57,529-707, src/cell_connectivity.cpp:35.
// The library source is absent from /root/reference; local vertex/face/cell ORDER inside a
// tet is this restatement's own (documented in DESIGN.md "per-tet complex conventions").
// PARITY UNPINNED at this boundary except in aggregate through the reference's own
// golden tests (tests/test_implicit_networks.cpp), which are replayed against this code.
#pragma once
#include <array>
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <vector>

namespace simplicial_arrangement {

template <typename Scalar, int DIM>
using Plane = std::array<Scalar, DIM + 1>; // values at the DIM+1 simplex vertices
template <typename Scalar, int DIM>
using Material = std::array<Scalar, DIM + 1>;

template <int DIM>
struct Arrangement
{
    static constexpr size_t None = std::numeric_limits<size_t>::max();
    // vertex = DIM plane ids (0..DIM: simplex faces b_i = 0; DIM+1+j: input plane j)
    std::vector<std::array<size_t, DIM>> vertices;
    struct Face
    {
        std::vector<size_t> vertices; // closed loop, CCW seen from the plane's positive side
        size_t supporting_plane = None;
        size_t positive_cell = None;
        size_t negative_cell = None;
    };
    std::vector<Face> faces;
    struct Cell
    {
        std::vector<size_t> faces;
    };
    std::vector<Cell> cells;
    // all three empty when no two planes coincide
    std::vector<size_t> unique_plane_indices; // plane id -> group id
    std::vector<std::vector<size_t>> unique_planes; // group id -> plane ids
    std::vector<bool> unique_plane_orientations; // plane id -> same orientation as group's first
};

template <int DIM>
struct MaterialInterface
{
    static constexpr size_t None = std::numeric_limits<size_t>::max();
    // vertex = DIM+1 material ids (0..DIM: simplex-boundary pseudo materials)
    std::vector<std::array<size_t, DIM + 1>> vertices;
    struct Face
    {
        std::vector<size_t> vertices;
        size_t positive_material_label = None;
        size_t negative_material_label = None;
    };
    std::vector<Face> faces;
    struct Cell
    {
        std::vector<size_t> faces;
        size_t material_label = None;
    };
    std::vector<Cell> cells;
    std::vector<size_t> unique_material_indices; // material id -> group id
    std::vector<std::vector<size_t>> unique_materials; // group id -> material ids
};

Arrangement<3> compute_arrangement(const std::vector<Plane<double, 3>>& planes);
MaterialInterface<3> compute_material_interface(const std::vector<Material<double, 3>>& materials);

// Statistics of the oracle engine (not part of upstream): calls that took the table path /
// the general path / needed the exact-arithmetic fallback.
struct EngineStats
{
    size_t lookups = 0, general = 0, exact_fallbacks = 0;
};
EngineStats& engine_stats();

} // namespace simplicial_arrangement
