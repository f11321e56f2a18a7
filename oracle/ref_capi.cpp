// ORACLE (test infrastructure, NOT product code).
//
// "Hybrid reference": the reference's OWN sources (/root/reference/src/{implicit_arrangement,
// material_interface,csg,extract_mesh,mesh_connectivity,pair_faces,topo_ray_shooting,
// cell_connectivity}.cpp) compiled in place (never copied) against shim headers and the restated
// per-tet engine (oracle/sa), exposed to Python through a flat C interface.  Used to (1) replay
// the reference's golden tests (tests/test_implicit_networks.cpp) through the restated engine,
// (2) validate the port's filter/extract/xyz against the reference's actual code, (3) time the
// reference's CPU path (bench.py --impl reference, cpu_baseline.kind = "reference").
#include <simplicial_arrangement/lookup_table.h>
#include <simplicial_arrangement/simplicial_arrangement.h>

#include "csg.h"
#include "extract_mesh.h"
#include "implicit_arrangement.h"
#include "material_interface.h"

#include "result_bag.h"
#ifndef RIN_GPU_DROPIN
#include "cell_connectivity.h"
#include "../robust-implicit-surface-networks_b200/host/cell_graph.h"
#endif
#ifdef RIN_GPU_DROPIN
#include "../robust-implicit-surface-networks_b200/host/rin_host.h"
#endif

#include <functional>
#include <iostream>

namespace {
struct CoutMute
{
    std::streambuf* old;
    explicit CoutMute(bool on) : old(nullptr)
    {
        if (on) old = std::cout.rdbuf(nullptr);
    }
    ~CoutMute()
    {
        if (old) {
            std::cout.rdbuf(old);
            std::cout.clear();
        }
    }
};

void pack_crs(ResultBag* bag, const std::string& name, const std::vector<std::vector<size_t>>& v)
{
    auto& off = bag->i64[name + "_offsets"];
    auto& dat = bag->i64[name];
    off.push_back(0);
    for (auto& l : v) {
        for (auto x : l) dat.push_back(x == Mesh_None ? -1 : int64_t(x));
        off.push_back(int64_t(dat.size()));
    }
}

void pack_mesh(ResultBag* bag, const std::vector<std::array<double, 3>>& pts,
    const std::vector<PolygonFace>& faces)
{
    auto& xyz = bag->f64["vert_xyz"];
    for (auto& p : pts) xyz.insert(xyz.end(), p.begin(), p.end());
    auto& foff = bag->i64["face_offsets"];
    auto& fv = bag->i64["face_verts"];
    auto& ftoff = bag->i64["face_tet_offsets"];
    auto& ft = bag->i64["face_tets"];
    auto& ff = bag->i64["face_funcs"];
    foff.push_back(0);
    ftoff.push_back(0);
    for (auto& f : faces) {
        for (auto v : f.vert_indices) fv.push_back(int64_t(v));
        foff.push_back(int64_t(fv.size()));
        for (auto& p : f.tet_face_indices) {
            ft.push_back(int64_t(p.first));
            ft.push_back(int64_t(p.second));
        }
        ftoff.push_back(int64_t(ft.size() / 2));
        ff.push_back(int64_t(f.func_index.first));
        ff.push_back(int64_t(f.func_index.second));
    }
}

void pack_labels(ResultBag* bag, const std::vector<std::string>& tl, const std::vector<double>& t,
    const std::vector<std::string>& sl, const std::vector<size_t>& s)
{
    std::string names;
    for (auto& l : tl) names += l + "\n";
    bag->error = ""; // not an error; labels travel in i64/f64 + a joined string below
    bag->f64["timings"] = t;
    auto& st = bag->i64["stats"];
    for (auto x : s) st.push_back(int64_t(x));
    bag->i64["timing_label_bytes"].assign(names.begin(), names.end());
    std::string sn;
    for (auto& l : sl) sn += l + "\n";
    bag->i64["stats_label_bytes"].assign(sn.begin(), sn.end());
}
} // namespace

extern "C" {

// flags: bit0 robust_test, bit1 use_lookup, bit2 use_secondary_lookup, bit3 use_topo_ray_shooting,
// bit4 quiet (mute std::cout)
void* ref_ia_run(const double* pts_in, uint64_t V, const uint64_t* tets_in, uint64_t T,
    const double* vals, uint32_t F, uint32_t flags)
{
    auto* bag = new ResultBag;
    CoutMute mute(flags & 16);
    std::vector<std::array<double, 3>> pts(V);
    for (uint64_t i = 0; i < V; ++i) pts[i] = {pts_in[3 * i], pts_in[3 * i + 1], pts_in[3 * i + 2]};
    std::vector<std::array<size_t, 4>> tets(T);
    for (uint64_t i = 0; i < T; ++i)
        tets[i] = {tets_in[4 * i], tets_in[4 * i + 1], tets_in[4 * i + 2], tets_in[4 * i + 3]};
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> fv(V, F);
    std::copy(vals, vals + V * F, fv.data());
    bool use_lookup = flags & 2;
    if (use_lookup) { // app/implicit_arrangement.cpp:29-42
        simplicial_arrangement::load_lookup_table(simplicial_arrangement::ARRANGEMENT);
        simplicial_arrangement::enable_lookup_table();
    } else
        simplicial_arrangement::disable_lookup_table();
    std::vector<std::array<double, 3>> iso_pts;
    std::vector<PolygonFace> iso_faces;
    std::vector<std::vector<size_t>> patches, chains, nme, shells, cells;
    std::vector<size_t> patch_label;
    std::vector<Edge> edges;
    std::vector<std::vector<bool>> cell_label;
    std::vector<std::string> tl, sl;
    std::vector<double> tm;
    std::vector<size_t> st;
    bool ok = false;
    try {
        ok = implicit_arrangement(flags & 1, use_lookup, flags & 4, flags & 8, pts, tets, fv, iso_pts,
            iso_faces, patches, patch_label, edges, chains, nme, shells, cells, cell_label, tl, tm, sl, st);
    } catch (std::exception& e) {
        bag->i64["threw"] = {1};
        pack_labels(bag, tl, tm, sl, st);
        pack_mesh(bag, iso_pts, iso_faces); // the hot-path outputs exist before host topology throws
        bag->error = e.what();
        return bag;
    }
    pack_labels(bag, tl, tm, sl, st);
    bag->i64["success"] = {ok ? 1 : 0};
    pack_mesh(bag, iso_pts, iso_faces);
    pack_crs(bag, "patches", patches);
    pack_crs(bag, "chains", chains);
    pack_crs(bag, "non_manifold_edges_of_vert", nme);
    pack_crs(bag, "shells", shells);
    pack_crs(bag, "cells", cells);
    auto& pl = bag->i64["patch_function_label"];
    for (auto x : patch_label) pl.push_back(int64_t(x));
#ifdef RIN_GPU_DROPIN
    bag->i64["complexes_fetched"].push_back(int64_t(rin_host::complexes_fetched()));
#endif
    auto& ed = bag->i64["edges"];
    auto& ef = bag->i64["edge_faces"];
    auto& efo = bag->i64["edge_faces_offsets"];
    efo.push_back(0);
    for (auto& e : edges) {
        ed.push_back(int64_t(e.v1));
        ed.push_back(int64_t(e.v2));
        for (auto& q : e.face_edge_indices) {
            ef.push_back(int64_t(q.first));
            ef.push_back(int64_t(q.second));
        }
        efo.push_back(int64_t(ef.size() / 2));
    }
    auto& cl = bag->i64["cell_function_label"]; // cells x F, row-major 0/1
    for (auto& row : cell_label)
        for (bool b : row) cl.push_back(b ? 1 : 0);
    return bag;
}

#ifndef RIN_GPU_DROPIN // the drop-in library does not link the reference's extraction
extern "C++" {
namespace {
// The product's simplicial-cell graph (host/cell_graph.h) against the reference's own build_simplicial_cell_adjacency
// + compute_simplicial_cell_connected_components (src/cell_connectivity.cpp) on the same inputs.  The adjacency only
// looks patches and shells up, so synthetic patch / shell numberings exercise it as well as real ones.
// Returns {adjacency arrays equal, components equal as shell sets, #simplicial cells, #components}.
template <typename Complex, typename PositiveSide>
std::vector<int64_t> check_cell_graph(const std::vector<std::array<size_t, 4>>& tets, const std::vector<Complex>& cut_results,
    const std::vector<size_t>& cut_result_index, const std::vector<long long>& gv, const std::vector<size_t>& gv_start,
    const std::vector<size_t>& ff, const std::vector<size_t>& ff_start, size_t n_faces, PositiveSide side)
{
    std::vector<size_t> patch_of_face(n_faces), shell_of_half_patch;
    size_t n_patches = 0;
    for (size_t f = 0; f < n_faces; ++f) {
        patch_of_face[f] = (f * 2654435761u) % (n_faces / 3 + 1);
        n_patches = std::max(n_patches, patch_of_face[f] + 1);
    }
    for (size_t h = 0; h < 2 * n_patches; ++h) shell_of_half_patch.push_back((h * 40503u) % (n_patches / 2 + 2));
    std::vector<std::pair<size_t, size_t>> tc_ref, tc_own;
    std::vector<long long> info_ref, info_own;
    std::vector<size_t> start_ref, start_own;
    build_simplicial_cell_adjacency(tets, cut_results, cut_result_index, gv, gv_start, ff, ff_start, patch_of_face,
        shell_of_half_patch, tc_ref, info_ref, start_ref);
    rin_host::simplicial_cell_adjacency(tets, cut_results, cut_result_index, gv, gv_start, ff, ff_start, patch_of_face,
        shell_of_half_patch, side, tc_own, info_own, start_own);
    std::vector<std::vector<size_t>> cells_ref, cells_own;
    compute_simplicial_cell_connected_components(tc_ref, info_ref, start_ref, cells_ref);
    rin_host::simplicial_cell_components(tc_own, info_own, start_own, cells_own);
    for (auto& c : cells_ref) std::sort(c.begin(), c.end()); // the reference's order inside a cell is a hash set's
    return {int64_t(tc_ref == tc_own && info_ref == info_own && start_ref == start_own), int64_t(cells_ref == cells_own),
        int64_t(tc_own.size()), int64_t(cells_own.size())};
}
} // namespace
} // extern "C++"
// The reference's cell-grouping extraction (second extract_iso_mesh overload, src/extract_mesh.cpp:268-566)
// called directly: its inputs are produced here the way implicit_arrangement() produces them (active
// function lists per tet, one arrangement per active tet), its maps are returned for the parity test of
// rin_tet_maps.
void* ref_ia_cellgroup_maps(const uint64_t* tets_in, uint64_t T, uint64_t V, const double* vals, uint32_t F)
{
    auto* bag = new ResultBag;
    using simplicial_arrangement::Arrangement;
    std::vector<std::array<size_t, 4>> tets(T);
    for (uint64_t i = 0; i < T; ++i)
        tets[i] = {tets_in[4 * i], tets_in[4 * i + 1], tets_in[4 * i + 2], tets_in[4 * i + 3]};
    simplicial_arrangement::disable_lookup_table();
    std::vector<size_t> func_in_tet, start_index_of_tet(1, 0), cut_result_index(T, Arrangement<3>::None);
    std::vector<Arrangement<3>> cut_results;
    size_t n1 = 0, n2 = 0, nm = 0;
    try {
        for (uint64_t i = 0; i < T; ++i) {
            std::vector<simplicial_arrangement::Plane<double, 3>> planes;
            for (uint32_t f = 0; f < F; ++f) {
                int pos = 0, neg = 0;
                for (int c = 0; c < 4; ++c) {
                    const double x = vals[tets[i][c] * F + f];
                    pos += x > 0;
                    neg += x < 0;
                }
                if (pos == 4 || neg == 4) continue;
                func_in_tet.push_back(f);
                planes.push_back({vals[tets[i][0] * F + f], vals[tets[i][1] * F + f], vals[tets[i][2] * F + f],
                    vals[tets[i][3] * F + f]});
            }
            start_index_of_tet.push_back(func_in_tet.size());
            if (planes.empty()) continue;
            cut_result_index[i] = cut_results.size();
            cut_results.emplace_back(simplicial_arrangement::compute_arrangement(planes));
            (planes.size() == 1 ? n1 : planes.size() == 2 ? n2 : nm)++;
        }
        std::vector<IsoVert> iso_verts;
        std::vector<PolygonFace> iso_faces;
        std::vector<long long> gv;
        std::vector<size_t> gv_start, ff, ff_start;
        extract_iso_mesh(n1, n2, nm, cut_results, cut_result_index, func_in_tet, start_index_of_tet, tets, iso_verts,
            iso_faces, gv, gv_start, ff, ff_start);
        auto& a = bag->i64["global_vId_of_tet_vert"];
        for (auto x : gv) a.push_back(int64_t(x));
        auto& b = bag->i64["global_vId_start_index_of_tet"];
        for (auto x : gv_start) b.push_back(int64_t(x));
        auto& c = bag->i64["iso_fId_of_tet_face"];
        for (auto x : ff) c.push_back(x == Arrangement<3>::None ? -1 : int64_t(x));
        auto& d = bag->i64["iso_fId_start_index_of_tet"];
        for (auto x : ff_start) d.push_back(int64_t(x));
        bag->i64["counts"] = {int64_t(iso_verts.size()), int64_t(iso_faces.size())};
        bag->i64["cell_graph"] = check_cell_graph(tets, cut_results, cut_result_index, gv, gv_start, ff, ff_start,
            iso_faces.size(), [](const Arrangement<3>& cx, size_t f, size_t cell) { return cx.faces[f].positive_cell == cell; });
    } catch (std::exception& e) {
        bag->error = e.what();
    }
    (void)V;
    return bag;
}

// Material-interface analogue (second extract_MI_mesh overload, src/extract_mesh.cpp:988-1443); the active
// material lists (material_in_tet / start_index_of_tet, src/material_interface.cpp:99-152) are passed in.
void* ref_mi_cellgroup_maps(const uint64_t* tets_in, uint64_t T, const double* vals, uint32_t F,
    const int64_t* material_in_tet_in, uint64_t n_mit, const int64_t* start_in)
{
    auto* bag = new ResultBag;
    using simplicial_arrangement::MaterialInterface;
    std::vector<std::array<size_t, 4>> tets(T);
    for (uint64_t i = 0; i < T; ++i)
        tets[i] = {tets_in[4 * i], tets_in[4 * i + 1], tets_in[4 * i + 2], tets_in[4 * i + 3]};
    simplicial_arrangement::disable_lookup_table();
    std::vector<size_t> material_in_tet(material_in_tet_in, material_in_tet_in + n_mit),
        start_index_of_tet(start_in, start_in + T + 1), cut_result_index(T, MaterialInterface<3>::None);
    std::vector<MaterialInterface<3>> cut_results;
    size_t n2 = 0, n3 = 0, nm = 0;
    try {
        for (uint64_t i = 0; i < T; ++i) {
            const size_t k = start_index_of_tet[i + 1] - start_index_of_tet[i];
            if (k == 0) continue;
            std::vector<simplicial_arrangement::Material<double, 3>> mats;
            for (size_t j = 0; j < k; ++j) {
                const size_t f = material_in_tet[start_index_of_tet[i] + j];
                mats.push_back({vals[tets[i][0] * F + f], vals[tets[i][1] * F + f], vals[tets[i][2] * F + f],
                    vals[tets[i][3] * F + f]});
            }
            cut_result_index[i] = cut_results.size();
            cut_results.emplace_back(simplicial_arrangement::compute_material_interface(mats));
            (k == 2 ? n2 : k == 3 ? n3 : nm)++;
        }
        std::vector<MI_Vert> verts;
        std::vector<PolygonFace> faces;
        std::vector<long long> gv;
        std::vector<size_t> gv_start, ff, ff_start;
        extract_MI_mesh(n2, n3, nm, cut_results, cut_result_index, material_in_tet, start_index_of_tet, tets, verts,
            faces, gv, gv_start, ff, ff_start);
        auto& a = bag->i64["global_vId_of_tet_vert"];
        for (auto x : gv) a.push_back(int64_t(x));
        auto& b = bag->i64["global_vId_start_index_of_tet"];
        for (auto x : gv_start) b.push_back(int64_t(x));
        auto& c = bag->i64["iso_fId_of_tet_face"];
        for (auto x : ff) c.push_back(x == MaterialInterface<3>::None ? -1 : int64_t(x));
        auto& d = bag->i64["iso_fId_start_index_of_tet"];
        for (auto x : ff_start) d.push_back(int64_t(x));
        bag->i64["counts"] = {int64_t(verts.size()), int64_t(faces.size())};
        bag->i64["cell_graph"] = check_cell_graph(tets, cut_results, cut_result_index, gv, gv_start, ff, ff_start,
            faces.size(), [](const MaterialInterface<3>& cx, size_t f, size_t cell) {
                return cx.faces[f].positive_material_label == cx.cells[cell].material_label;
            });
    } catch (std::exception& e) {
        bag->error = e.what();
    }
    return bag;
}
#endif

void* ref_mi_run(const double* pts_in, uint64_t V, const uint64_t* tets_in, uint64_t T,
    const double* vals, uint32_t F, uint32_t flags)
{
    auto* bag = new ResultBag;
    CoutMute mute(flags & 16);
    std::vector<std::array<double, 3>> pts(V);
    for (uint64_t i = 0; i < V; ++i) pts[i] = {pts_in[3 * i], pts_in[3 * i + 1], pts_in[3 * i + 2]};
    std::vector<std::array<size_t, 4>> tets(T);
    for (uint64_t i = 0; i < T; ++i)
        tets[i] = {tets_in[4 * i], tets_in[4 * i + 1], tets_in[4 * i + 2], tets_in[4 * i + 3]};
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> fv(V, F);
    std::copy(vals, vals + V * F, fv.data());
    bool use_lookup = flags & 2;
    if (use_lookup) { // app/material_interface.cpp:32-45
        simplicial_arrangement::load_lookup_table(simplicial_arrangement::MATERIAL_INTERFACE);
        simplicial_arrangement::enable_lookup_table();
    } else
        simplicial_arrangement::disable_lookup_table();
    std::vector<std::array<double, 3>> mi_pts;
    std::vector<PolygonFace> mi_faces;
    std::vector<std::vector<size_t>> patches, chains, nme, shells, cells;
    std::vector<std::pair<size_t, size_t>> patch_label;
    std::vector<Edge> edges;
    std::vector<size_t> cell_label;
    std::vector<std::string> tl, sl;
    std::vector<double> tm;
    std::vector<size_t> st;
    bool ok = false;
    try {
        ok = material_interface(flags & 1, use_lookup, flags & 4, flags & 8, pts, tets, fv, mi_pts,
            mi_faces, patches, patch_label, edges, chains, nme, shells, cells, cell_label, tl, tm, sl, st);
    } catch (std::exception& e) {
        bag->i64["threw"] = {1};
        pack_labels(bag, tl, tm, sl, st);
        pack_mesh(bag, mi_pts, mi_faces);
        bag->error = e.what();
        return bag;
    }
    pack_labels(bag, tl, tm, sl, st);
    bag->i64["success"] = {ok ? 1 : 0};
    pack_mesh(bag, mi_pts, mi_faces);
    pack_crs(bag, "patches", patches);
    pack_crs(bag, "chains", chains);
    pack_crs(bag, "non_manifold_edges_of_vert", nme);
    pack_crs(bag, "shells", shells);
    pack_crs(bag, "cells", cells);
    auto& pl = bag->i64["patch_function_label"];
    for (auto& x : patch_label) {
        pl.push_back(int64_t(x.first));
        pl.push_back(int64_t(x.second));
    }
#ifdef RIN_GPU_DROPIN
    bag->i64["complexes_fetched"].push_back(int64_t(rin_host::complexes_fetched()));
#endif
    auto& ed = bag->i64["edges"];
    auto& ef = bag->i64["edge_faces"];
    auto& efo = bag->i64["edge_faces_offsets"];
    efo.push_back(0);
    for (auto& e : edges) {
        ed.push_back(int64_t(e.v1));
        ed.push_back(int64_t(e.v2));
        for (auto& q : e.face_edge_indices) {
            ef.push_back(int64_t(q.first));
            ef.push_back(int64_t(q.second));
        }
        efo.push_back(int64_t(ef.size() / 2));
    }
    auto& cl = bag->i64["cell_function_label"];
    for (auto x : cell_label) cl.push_back(x == Mesh_None ? -1 : int64_t(x));
    return bag;
}

// csg() (src/csg.h:46-72) with one of the boolean expressions of the reference's CSG tests
// (tests/test_implicit_networks.cpp:865,910,955,1000): expr 0: f0 & !f0, 1: f0 | (f1 & f2),
// 2: f0 & !(f1 & f2), 3: f1 | f2.
void* ref_csg_run(const double* pts_in, uint64_t V, const uint64_t* tets_in, uint64_t T, const double* vals,
    uint32_t F, uint32_t flags, int expr, int positive_inside)
{
    auto* bag = new ResultBag;
    CoutMute mute(flags & 16);
    std::vector<std::array<double, 3>> pts(V);
    for (uint64_t i = 0; i < V; ++i) pts[i] = {pts_in[3 * i], pts_in[3 * i + 1], pts_in[3 * i + 2]};
    std::vector<std::array<size_t, 4>> tets(T);
    for (uint64_t i = 0; i < T; ++i)
        tets[i] = {tets_in[4 * i], tets_in[4 * i + 1], tets_in[4 * i + 2], tets_in[4 * i + 3]};
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> fv(V, F);
    std::copy(vals, vals + V * F, fv.data());
    bool use_lookup = flags & 2;
    if (use_lookup) {
        simplicial_arrangement::load_lookup_table(simplicial_arrangement::ARRANGEMENT);
        simplicial_arrangement::enable_lookup_table();
    } else
        simplicial_arrangement::disable_lookup_table();
    std::function<bool(std::vector<bool>)> lambda;
    switch (expr) {
    case 0: lambda = [](std::vector<bool> c) { return c[0] && !c[0]; }; break;
    case 1: lambda = [](std::vector<bool> c) { return c[0] || (c[1] && c[2]); }; break;
    case 2: lambda = [](std::vector<bool> c) { return c[0] && !(c[1] && c[2]); }; break;
    default: lambda = [](std::vector<bool> c) { return c[1] || c[2]; }; break;
    }
    std::vector<std::array<double, 3>> iso_pts;
    std::vector<PolygonFace> iso_faces;
    std::vector<std::vector<size_t>> patches, chains, nme, shells, cells;
    std::vector<size_t> patch_label;
    std::vector<bool> patch_sign;
    std::vector<Edge> edges;
    std::vector<std::vector<bool>> cell_label;
    std::vector<std::string> tl, sl;
    std::vector<double> tm;
    std::vector<size_t> st;
    bool ok = false;
    try {
        ok = csg(flags & 1, use_lookup, flags & 4, flags & 8, positive_inside != 0, pts, tets, fv, lambda, iso_pts,
            iso_faces, patches, patch_label, patch_sign, edges, chains, nme, shells, cells, cell_label, tl, tm, sl, st);
    } catch (std::exception& e) {
        bag->i64["threw"] = {1};
        bag->error = e.what();
        return bag;
    }
    bag->i64["success"] = {ok ? 1 : 0};
    pack_mesh(bag, iso_pts, iso_faces);
    pack_crs(bag, "patches", patches);
    pack_crs(bag, "chains", chains);
    pack_crs(bag, "non_manifold_edges_of_vert", nme);
    auto& ps = bag->i64["patch_sign_label"];
    for (bool b : patch_sign) ps.push_back(b ? 1 : 0);
    return bag;
}

const int64_t* ref_i64(void* h, const char* name, uint64_t* n)
{
    auto* b = static_cast<ResultBag*>(h);
    auto it = b->i64.find(name);
    if (it == b->i64.end()) {
        *n = 0;
        return nullptr;
    }
    *n = it->second.size();
    return it->second.data();
}
const double* ref_f64(void* h, const char* name, uint64_t* n)
{
    auto* b = static_cast<ResultBag*>(h);
    auto it = b->f64.find(name);
    if (it == b->f64.end()) {
        *n = 0;
        return nullptr;
    }
    *n = it->second.size();
    return it->second.data();
}
const char* ref_error(void* h)
{
    return static_cast<ResultBag*>(h)->error.c_str();
}
void ref_free(void* h)
{
    delete static_cast<ResultBag*>(h);
}
}
