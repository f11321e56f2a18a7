// Simplicial-cell graph of the cell-grouping path (useTopoRayShooting = false), SURVEY 8(f) N4:
//   simplicial_cell_adjacency  = build_simplicial_cell_adjacency (src/cell_connectivity.cpp:15-163 implicit
//                                arrangement, :165-312 material interface)
//   simplicial_cell_components = compute_simplicial_cell_connected_components (:314-372)
// Host side by north_star; written for this repository (no reference code), header-only so that it serves the
// reference's complex types (simplicial_arrangement::Arrangement<3> / MaterialInterface<3>) and rin_host::TetComplex
// alike: `on_positive_side(complex, face, cell)` tells whether the cell is on the positive side of the face
// (IA: faces[f].positive_cell == cell, :78; MI: faces[f].positive_material_label == cells[cell].material_label, :227).
//
// Outputs, exactly the reference's encoding: one simplicial cell per cell of every active tet's complex and one per
// empty tet, in tet order; per half-face (cell's face list order; the four faces opposite corner 0..3 for an empty
// tet): -(shell id) - 1 for a face on the surface, the neighbouring simplicial cell for a face shared with another
// cell, LLONG_MAX for a face on the mesh boundary.  The reference pairs half-faces with a hash map keyed by
// (smallest, second smallest, largest) global vertex id (compute_iso_face_key, src/extract_mesh.h:68-92) with
// insert-then-erase; here the half-faces are sorted by that key (stable, so creation order survives) and paired two
// by two, which gives the same pairs, also when more than two half-faces carry one key.
#pragma once
#include <algorithm>
#include <array>
#include <cstddef>
#include <limits>
#include <utility>
#include <vector>

namespace rin_host {

inline std::array<long long, 3> half_face_key(const long long* v, size_t n)
{
    long long lo = v[0], hi = v[0];
    size_t lo_at = 0;
    for (size_t i = 1; i < n; ++i) {
        if (v[i] < lo) {
            lo = v[i];
            lo_at = i;
        } else if (v[i] > hi)
            hi = v[i];
    }
    long long second = hi + 1;
    for (size_t i = 0; i < n; ++i)
        if (i != lo_at && v[i] < second) second = v[i];
    return {lo, second, hi};
}

template <typename Complex, typename PositiveSide>
void simplicial_cell_adjacency(const std::vector<std::array<size_t, 4>>& tets, const std::vector<Complex>& cut_results,
    const std::vector<size_t>& cut_result_index, const std::vector<long long>& global_vId_of_tet_vert,
    const std::vector<size_t>& global_vId_start_index_of_tet, const std::vector<size_t>& fId_of_tet_face,
    const std::vector<size_t>& fId_start_index_of_tet, const std::vector<size_t>& patch_of_face,
    const std::vector<size_t>& shell_of_half_patch, PositiveSide on_positive_side,
    std::vector<std::pair<size_t, size_t>>& tet_cell_of_simp_cell, std::vector<long long>& simp_half_face_info,
    std::vector<size_t>& simp_hFace_start_index)
{
    constexpr size_t None = std::numeric_limits<size_t>::max();
    constexpr long long Boundary = std::numeric_limits<long long>::max();
    struct Open
    {
        std::array<long long, 3> key;
        size_t slot; // index into simp_half_face_info
        size_t cell;
    };
    std::vector<Open> open;
    simp_hFace_start_index.push_back(0);
    std::vector<long long> gv;
    for (size_t t = 0; t < tets.size(); ++t) {
        if (cut_result_index[t] == None) {
            // an empty tet is one cell; face c is opposite corner c, corners are encoded -(vertex) - 1
            const size_t cell = tet_cell_of_simp_cell.size();
            tet_cell_of_simp_cell.emplace_back(t, 0);
            for (int c = 0; c < 4; ++c) {
                long long v[3];
                for (int q = 0, w = 0; q < 4; ++q)
                    if (q != c) v[w++] = -(long long)tets[t][q] - 1;
                open.push_back({half_face_key(v, 3), simp_half_face_info.size(), cell});
                simp_half_face_info.push_back(Boundary);
            }
            simp_hFace_start_index.push_back(simp_half_face_info.size());
            continue;
        }
        const Complex& cx = cut_results[cut_result_index[t]];
        const size_t f0 = fId_start_index_of_tet[t], v0 = global_vId_start_index_of_tet[t];
        for (size_t j = 0; j < cx.cells.size(); ++j) {
            const size_t cell = tet_cell_of_simp_cell.size();
            tet_cell_of_simp_cell.emplace_back(t, j);
            for (size_t f : cx.cells[j].faces) {
                const size_t surface_face = fId_of_tet_face[f0 + f];
                if (surface_face != None) {
                    const size_t half_patch = 2 * patch_of_face[surface_face] + (on_positive_side(cx, f, j) ? 0 : 1);
                    simp_half_face_info.push_back(-(long long)shell_of_half_patch[half_patch] - 1);
                    continue;
                }
                gv.clear();
                for (size_t lv : cx.faces[f].vertices) gv.push_back(global_vId_of_tet_vert[v0 + lv]);
                open.push_back({half_face_key(gv.data(), gv.size()), simp_half_face_info.size(), cell});
                simp_half_face_info.push_back(Boundary);
            }
            simp_hFace_start_index.push_back(simp_half_face_info.size());
        }
    }
    std::stable_sort(open.begin(), open.end(), [](const Open& a, const Open& b) { return a.key < b.key; });
    for (size_t i = 0; i + 1 < open.size();) {
        if (open[i].key == open[i + 1].key) {
            simp_half_face_info[open[i].slot] = (long long)open[i + 1].cell;
            simp_half_face_info[open[i + 1].slot] = (long long)open[i].cell;
            i += 2;
        } else
            ++i;
    }
}

// Connected components of the simplicial-cell graph; a component (arrangement cell) is reported as the list of
// shells its cells touch.  Components appear in the order of their lowest cell.  The reference collects the shells
// in an absl::flat_hash_set, so their order inside a cell is unspecified there; here they ascend.
inline void simplicial_cell_components(const std::vector<std::pair<size_t, size_t>>& tet_cell_of_simp_cell,
    const std::vector<long long>& simp_half_face_info, const std::vector<size_t>& simp_hFace_start_index,
    std::vector<std::vector<size_t>>& arrangement_cells)
{
    constexpr long long Boundary = std::numeric_limits<long long>::max();
    const size_t n = tet_cell_of_simp_cell.size();
    std::vector<char> seen(n, 0);
    std::vector<size_t> frontier;
    for (size_t root = 0; root < n; ++root) {
        if (seen[root]) continue;
        std::vector<size_t> shells;
        frontier.assign(1, root);
        seen[root] = 1;
        for (size_t head = 0; head < frontier.size(); ++head) {
            const size_t c = frontier[head];
            for (size_t h = simp_hFace_start_index[c]; h < simp_hFace_start_index[c + 1]; ++h) {
                const long long x = simp_half_face_info[h];
                if (x < 0)
                    shells.push_back((size_t)(-x - 1));
                else if (x != Boundary && !seen[(size_t)x]) {
                    seen[(size_t)x] = 1;
                    frontier.push_back((size_t)x);
                }
            }
        }
        std::sort(shells.begin(), shells.end());
        shells.erase(std::unique(shells.begin(), shells.end()), shells.end());
        arrangement_cells.push_back(std::move(shells));
    }
}

} // namespace rin_host
