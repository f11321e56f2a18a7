// Config parser, tet-mesh loader and result writers of the drop-in: the reference's on-disk
// contract (/root/reference/src/io.h:12-91, src/io.cpp:19-93, 154-372, 374-588) without its third-party
// dependencies (nlohmann/json, ghc::filesystem, MshIO).  Same argument lists; the JSON files are byte-identical to
// what the reference writes through nlohmann::json (keys in lexicographic order, one line, `null` for arrays that
// received no element, shortest round-trip doubles in nlohmann's notation).
#pragma once
#include "mesh_types.h"

#include <array>
#include <string>
#include <vector>

namespace rin_host {

struct Config // src/io.h:12-28
{
    std::string tet_mesh_file, func_file, output_dir;
    bool use_lookup = true, use_secondary_lookup = true, use_topo_ray_shooting = true;
    size_t tet_mesh_resolution = 0;
    std::array<double, 3> tet_mesh_bbox_min{}, tet_mesh_bbox_max{};
};

// throws std::runtime_error("Config file does not exist!") like the reference; relative paths are resolved against
// the directory of the config file; tetMeshFile wins over gridResolution / gridBbox (src/io.cpp:36-48)
Config parse_config_file(const std::string& filename);

// [[ [x,y,z]... ], [ [a,b,c,d]... ]]  (src/io.cpp:65-93)
bool load_tet_mesh(const std::string& filename, std::vector<std::array<double, 3>>& pts,
    std::vector<std::array<size_t, 4>>& tets);

bool save_result(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<size_t>& patch_function_label, const std::vector<Edge>& edges,
    const std::vector<std::vector<size_t>>& chains, const std::vector<std::vector<size_t>>& non_manifold_edges_of_vert,
    const std::vector<std::vector<size_t>>& shells, const std::vector<std::vector<size_t>>& cells,
    const std::vector<std::vector<bool>>& cell_function_label);

bool save_result_MI(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<std::pair<size_t, size_t>>& patch_function_label, const std::vector<Edge>& edges,
    const std::vector<std::vector<size_t>>& chains, const std::vector<std::vector<size_t>>& non_manifold_edges_of_vert,
    const std::vector<std::vector<size_t>>& shells, const std::vector<std::vector<size_t>>& cells,
    const std::vector<size_t>& cell_function_label);

bool save_result_CSG(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<bool>& patch_sign_label, const std::vector<Edge>& edges,
    const std::vector<std::vector<size_t>>& chains, const std::vector<std::vector<size_t>>& non_manifold_edges_of_vert);

// <filename>_chains.msh, <filename>_patches.msh (attributes patch_id, polygon_id), <filename>_cells.msh (cell_id)
bool save_result_msh(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<Edge>& edges, const std::vector<std::vector<size_t>>& chains,
    const std::vector<std::vector<size_t>>& non_manifold_edges_of_vert, const std::vector<std::vector<size_t>>& shells,
    const std::vector<std::vector<size_t>>& cells);

bool save_timings(const std::string& filename, const std::vector<std::string>& timing_labels,
    const std::vector<double>& timings);
bool save_statistics(const std::string& filename, const std::vector<std::string>& stats_labels,
    const std::vector<size_t>& stats);

// the number notation of the JSON files (exposed for tests)
std::string json_number(double x);

} // namespace rin_host
