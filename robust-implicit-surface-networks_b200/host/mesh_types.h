// Output records of the hot path, field for field those of the reference
// (/root/reference/src/mesh.h:11-71) so that the reference's host stages consume them unchanged.
// When the reference's own mesh.h is already included (integration builds) these are not redefined.
#ifndef ROBUST_IMPLICIT_NETWORKS_MESH_H
#define ROBUST_IMPLICIT_NETWORKS_MESH_H
#include <array>
#include <cstddef>
#include <limits>
#include <utility>
#include <vector>

static constexpr size_t Mesh_None = std::numeric_limits<size_t>::max();

struct PolygonFace
{
    std::vector<size_t> vert_indices;
    std::vector<std::pair<size_t, size_t>> tet_face_indices; // (tet, local face id)
    std::pair<size_t, size_t> func_index;                    // IA: first only; MI: (positive, negative) material
};

struct IsoVert
{
    size_t tet_index;
    size_t tet_vert_index;
    size_t simplex_size; // 1 point, 2 edge, 3 triangle, 4 tetrahedron
    std::array<size_t, 4> simplex_vert_indices;
    std::array<size_t, 3> func_indices = {Mesh_None, Mesh_None, Mesh_None};
};

struct MI_Vert
{
    size_t tet_index;
    size_t tet_vert_index;
    size_t simplex_size;
    std::array<size_t, 4> simplex_vert_indices;
    std::array<size_t, 4> material_indices;
};

struct Edge
{
    size_t v1;
    size_t v2;
    std::vector<std::pair<size_t, size_t>> face_edge_indices;
};
#endif
