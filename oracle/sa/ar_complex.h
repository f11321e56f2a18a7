// ORACLE (test infrastructure, NOT product code).  See ar_complex.cpp.
#pragma once
#include <simplicial_arrangement/simplicial_arrangement.h>

#include <array>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sa_oracle {

constexpr size_t NONE = SIZE_MAX;

struct ARComplex
{
    struct Edge
    {
        size_t v0, v1; // end points
        size_t p0, p1; // the two planes containing the edge (ascending)
    };
    struct Face
    {
        std::vector<size_t> verts; // loop; edges[k] joins verts[k] and verts[k+1]
        std::vector<size_t> edges;
        size_t plane = NONE, pos_cell = NONE, neg_cell = NONE;
    };
    struct Cell
    {
        std::vector<size_t> faces;
    };
    std::vector<std::array<size_t, 3>> vertices;
    std::vector<Edge> edges;
    std::vector<Face> faces;
    std::vector<Cell> cells;

    void init_simplex();
    // inserts plane `pid`; returns the id of an existing plane it coincides with, or NONE
    size_t add_plane(const std::vector<std::array<double, 4>>& planes, size_t pid);
};

int ar_vertex_orientation(
    const std::vector<std::array<double, 4>>& planes, const std::array<size_t, 3>& v, const double* q);

simplicial_arrangement::Arrangement<3> compute_arrangement_general(
    const std::vector<std::array<double, 4>>& input);
simplicial_arrangement::MaterialInterface<3> compute_material_interface_general(
    const std::vector<std::array<double, 4>>& input);

simplicial_arrangement::EngineStats& stats();

// lookup keys (shared by the table generator and the lookup): -1 when a sign is zero / degenerate
int ia_key_1(const double* p0);
int ia_key_2(const double* p0, const double* p1); // outer (8 bits) | inner (6 bits) << 8

} // namespace sa_oracle
