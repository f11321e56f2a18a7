#include "rin_host.h"

#include "../../include/rin_b200.h"

#include <cstdlib>
#include <iostream>
#include <mutex>

namespace rin_host {
namespace {

// One engine context per process, like the reference's process-global lookup tables (the entry points are not
// re-entrant there either): created on first use on the device RIN_DEVICE names (default 0), destroyed at
// exit.  Creation is guarded; calls that follow one another on different threads must be serialised by the caller.
rin_ctx* g_ctx = nullptr;
std::mutex g_ctx_mutex;
size_t g_complexes_fetched = 0; // per-tet complexes fetched since the last hot-path call

bool ensure_ctx(std::string& err)
{
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    if (g_ctx) return true;
    const char* dev = std::getenv("RIN_DEVICE");
    if (rin_create(dev ? std::atoi(dev) : 0, &g_ctx) != RIN_OK) {
        err = rin_last_error();
        g_ctx = nullptr;
        return false;
    }
    std::atexit([] {
        if (g_ctx) rin_destroy(g_ctx);
        g_ctx = nullptr;
    });
    return true;
}

size_t widen(uint32_t x)
{
    return x == 0xffffffffu ? Mesh_None : size_t(x);
}

struct Downloaded
{
    rin_counts n{};
    std::vector<uint32_t> vt, vs, vf, fo, fv, fto, ft, ff;
    std::vector<uint8_t> vl, vz;
};

bool run_and_download(int mode, uint32_t flags, const std::vector<std::array<double, 3>>& pts,
    const std::vector<std::array<size_t, 4>>& tets, const double* funcVals, size_t n_func,
    std::vector<std::array<double, 3>>& out_pts, Downloaded& d, HotPathOutput& out)
{
    if (!ensure_ctx(out.error)) return false;
    g_complexes_fetched = 0;
    static_assert(sizeof(size_t) == 8, "the reference's tets are 64-bit indices");
    if (rin_run_host(g_ctx, mode, flags, pts.empty() ? nullptr : pts[0].data(), pts.size(),
            tets.empty() ? nullptr : tets[0].data(), tets.size(), 8, funcVals, (uint32_t)n_func, &d.n) != RIN_OK) {
        out.error = rin_last_error();
        return false;
    }
    const rin_counts& n = d.n;
    d.vt.resize(n.num_verts);
    d.vl.resize(n.num_verts);
    d.vz.resize(n.num_verts);
    d.vs.resize(4 * n.num_verts);
    d.vf.resize(4 * n.num_verts);
    d.fo.resize(n.num_faces + 1);
    d.fv.resize(n.num_face_verts);
    d.fto.resize(n.num_faces + 1);
    d.ft.resize(2 * n.num_face_tets);
    d.ff.resize(2 * n.num_faces);
    out_pts.resize(n.num_verts);
    rin_mesh_out mo{d.vt.data(), d.vl.data(), d.vz.data(), d.vs.data(), d.vf.data(),
        out_pts.empty() ? nullptr : out_pts[0].data(), d.fo.data(), d.fv.data(), d.fto.data(), d.ft.data(),
        d.ff.data()};
    if (rin_download_mesh(g_ctx, &mo) != RIN_OK) {
        out.error = rin_last_error();
        return false;
    }
    std::vector<uint32_t> fit(n.num_active_funcs);
    std::vector<uint64_t> start(n.num_tets + 1);
    if (rin_download_active(g_ctx, fit.data(), start.data()) != RIN_OK) {
        out.error = rin_last_error();
        return false;
    }
    out.func_in_tet.assign(fit.begin(), fit.end());
    out.start_index_of_tet.assign(start.begin(), start.end());
    out.num_degenerate_vertex = n.num_degenerate_vertex;
    out.num_intersecting_tet = n.num_intersecting_tet;
    out.num_k1 = n.num_k1;
    out.num_k2 = n.num_k2;
    out.num_kmore = n.num_kmore;
    return true;
}

void fill_faces(const Downloaded& d, bool both_labels, std::vector<PolygonFace>& faces)
{
    const size_t nf = d.n.num_faces;
    faces.resize(nf);
    for (size_t f = 0; f < nf; ++f) {
        PolygonFace& F = faces[f];
        F.vert_indices.assign(d.fv.begin() + d.fo[f], d.fv.begin() + d.fo[f + 1]);
        for (uint32_t k = d.fto[f]; k < d.fto[f + 1]; ++k) F.tet_face_indices.emplace_back(d.ft[2 * k], d.ft[2 * k + 1]);
        F.func_index.first = widen(d.ff[2 * f]);
        F.func_index.second = both_labels ? widen(d.ff[2 * f + 1]) : F.func_index.second;
    }
}

void stage_times(bool mi, std::vector<std::string>& labels, std::vector<double>& timings)
{
    // device stage times folded into the reference's label set (src/implicit_arrangement.cpp:62-398,
    // src/material_interface.cpp:60-443), in seconds
    float ms[16] = {};
    const int k = rin_num_stages();
    rin_get_stage_times(g_ctx, ms, 16);
    auto s = [&](int i) { return i < k ? double(ms[i]) * 1e-3 : 0.0; };
    // stages: 0 eval+signs, 1 filter, 2 classify, 3 general, 4 count+scan, 5 emit, 6 dedup, 7 verts+xyz, 8 faces
    labels.emplace_back(mi ? "highest func" : "func signs");
    timings.push_back(s(0));
    labels.emplace_back("filter");
    timings.push_back(s(1));
    labels.emplace_back(mi ? "MI(other)" : "simp_arr(other)");
    timings.push_back(0.0);
    labels.emplace_back(mi ? "MI(2 func)" : "simp_arr(1 func)"); // table dispatch (IA: fused into the filter)
    timings.push_back(s(2));
    labels.emplace_back(mi ? "MI(3 func)" : "simp_arr(2 func)");
    timings.push_back(0.0);
    labels.emplace_back(mi ? "MI(>=4 func)" : "simp_arr(>=3 func)");
    timings.push_back(s(3));
    labels.emplace_back("extract mesh");
    timings.push_back(s(4) + s(5) + s(6) + s(8));
    labels.emplace_back("compute xyz");
    timings.push_back(s(7));
}

} // namespace

bool implicit_arrangement_hot(bool use_lookup, bool use_secondary_lookup,
    const std::vector<std::array<double, 3>>& pts, const std::vector<std::array<size_t, 4>>& tets,
    const double* funcVals, size_t n_func, bool negate, std::vector<std::array<double, 3>>& iso_pts,
    std::vector<PolygonFace>& iso_faces, std::vector<IsoVert>& iso_verts, HotPathOutput& out,
    std::vector<std::string>& timing_labels, std::vector<double>& timings, std::vector<std::string>& stats_labels,
    std::vector<size_t>& stats)
{
    if (!use_lookup) use_secondary_lookup = false; // src/implicit_arrangement.cpp:38-40
    const uint32_t flags = (use_lookup ? RIN_FLAG_USE_LOOKUP : 0u) |
                           (use_secondary_lookup ? RIN_FLAG_USE_SECONDARY_LOOKUP : 0u) | (negate ? RIN_FLAG_NEGATE : 0u);
    Downloaded d;
    if (!run_and_download(RIN_MODE_IA, flags, pts, tets, funcVals, n_func, iso_pts, d, out)) {
        std::cout << out.error << std::endl; // same contract as :302-305: message, return false
        return false;
    }
    const size_t nv = d.n.num_verts;
    iso_verts.resize(nv);
    for (size_t i = 0; i < nv; ++i) {
        IsoVert& v = iso_verts[i];
        v.tet_index = d.vt[i];
        v.tet_vert_index = d.vl[i];
        v.simplex_size = d.vz[i];
        for (int k = 0; k < 4; ++k) v.simplex_vert_indices[k] = widen(d.vs[4 * i + k]);
        for (int k = 0; k < 3; ++k) v.func_indices[k] = widen(d.vf[4 * i + k]);
    }
    fill_faces(d, false, iso_faces);
    stage_times(false, timing_labels, timings);
    const char* names[] = {"num_degenerate_vertex", "num_intersecting_tet", "num_1_func", "num_2_func", "num_more_func",
        "num_iso_verts", "num_iso_faces"};
    const size_t vals[] = {out.num_degenerate_vertex, out.num_intersecting_tet, out.num_k1, out.num_k2, out.num_kmore,
        nv, (size_t)d.n.num_faces};
    for (int i = 0; i < 7; ++i) {
        stats_labels.emplace_back(names[i]);
        stats.push_back(vals[i]);
    }
    return true;
}

bool material_interface_hot(bool use_lookup, bool use_secondary_lookup,
    const std::vector<std::array<double, 3>>& pts, const std::vector<std::array<size_t, 4>>& tets,
    const double* funcVals, size_t n_func, std::vector<std::array<double, 3>>& MI_pts,
    std::vector<PolygonFace>& MI_faces, std::vector<MI_Vert>& MI_verts, HotPathOutput& out,
    std::vector<std::string>& timing_labels, std::vector<double>& timings, std::vector<std::string>& stats_labels,
    std::vector<size_t>& stats)
{
    if (!use_lookup) use_secondary_lookup = false;
    const uint32_t flags = (use_lookup ? RIN_FLAG_USE_LOOKUP : 0u) | (use_secondary_lookup ? RIN_FLAG_USE_SECONDARY_LOOKUP : 0u);
    Downloaded d;
    if (!run_and_download(RIN_MODE_MI, flags, pts, tets, funcVals, n_func, MI_pts, d, out)) {
        std::cout << out.error << std::endl;
        return false;
    }
    const size_t nv = d.n.num_verts;
    MI_verts.resize(nv);
    for (size_t i = 0; i < nv; ++i) {
        MI_Vert& v = MI_verts[i];
        v.tet_index = d.vt[i];
        v.tet_vert_index = d.vl[i];
        v.simplex_size = d.vz[i];
        for (int k = 0; k < 4; ++k) {
            v.simplex_vert_indices[k] = widen(d.vs[4 * i + k]);
            v.material_indices[k] = widen(d.vf[4 * i + k]);
        }
    }
    fill_faces(d, true, MI_faces);
    stage_times(true, timing_labels, timings);
    const char* names[] = {"num_intersecting_tet", "num_2_func", "num_3_func", "num_more_func", "num_MI_verts",
        "num_MI_faces"};
    const size_t vals[] = {out.num_intersecting_tet, out.num_k1, out.num_k2, out.num_kmore, nv, (size_t)d.n.num_faces};
    for (int i = 0; i < 6; ++i) {
        stats_labels.emplace_back(names[i]);
        stats.push_back(vals[i]);
    }
    return true;
}

bool robust_test_verdict(int mode, std::string& error)
{
    if (!g_ctx) {
        error = "no hot-path run";
        return false;
    }
    uint32_t r[4] = {0, 0, 0, 0};
    if (rin_robust_test(g_ctx, mode, r) != RIN_OK) {
        error = rin_last_error();
        std::cout << error << std::endl;
        return false;
    }
    if (r[1]) {
        std::cout << "type 2 failure (crash in the normal order)." << std::endl;
        return false;
    }
    if (r[2]) {
        std::cout << "type 3 failure (crash in the reverse order)." << std::endl;
        return false;
    }
    if (r[0]) {
        std::cout << "type 1 failure (inconsistency)." << std::endl;
        return false;
    }
    std::cout << "success." << std::endl;
    return true;
}

size_t complexes_fetched()
{
    return g_complexes_fetched;
}

bool fetch_complexes(int mode, const std::vector<size_t>& tet_ids, std::vector<TetComplex>& out, std::string& error)
{
    if (!g_ctx) {
        error = "no hot-path run";
        return false;
    }
    g_complexes_fetched += tet_ids.size();
    std::vector<uint64_t> ids(tet_ids.begin(), tet_ids.end()), off(tet_ids.size() + 1);
    uint64_t nw = 0;
    if (rin_get_complexes(g_ctx, mode, 0, ids.data(), ids.size(), off.data(), nullptr, &nw) != RIN_OK) {
        error = rin_last_error();
        return false;
    }
    std::vector<uint32_t> w(nw ? nw : 1);
    if (rin_get_complexes(g_ctx, mode, 0, ids.data(), ids.size(), off.data(), w.data(), &nw) != RIN_OK) {
        error = rin_last_error();
        return false;
    }
    out.assign(tet_ids.size(), TetComplex{});
    const int vw = (mode == RIN_MODE_IA) ? 3 : 4;
    for (size_t i = 0; i < tet_ids.size(); ++i) {
        if (off[i + 1] == off[i]) continue;
        const uint32_t* p = w.data() + off[i];
        const uint32_t nv = p[0], nf = p[1], nc = p[2], nu = p[3];
        p += 4;
        TetComplex& cx = out[i];
        cx.vertices.resize(nv);
        for (uint32_t v = 0; v < nv; ++v) {
            cx.vertices[v] = {Mesh_None, Mesh_None, Mesh_None, Mesh_None};
            for (int k = 0; k < vw; ++k) cx.vertices[v][k] = *p++;
        }
        cx.faces.resize(nf);
        for (uint32_t f = 0; f < nf; ++f) {
            auto& F = cx.faces[f];
            F.a = widen(*p++);
            F.b = widen(*p++);
            if (mode == RIN_MODE_IA) F.c = widen(*p++);
            const uint32_t n = *p++;
            F.vertices.assign(p, p + n);
            p += n;
        }
        cx.cells.resize(nc);
        for (uint32_t c = 0; c < nc; ++c) {
            if (mode != RIN_MODE_IA) cx.cells[c].material_label = *p++;
            const uint32_t n = *p++;
            cx.cells[c].faces.assign(p, p + n);
            p += n;
        }
        if (nu) {
            const uint32_t np = *p++;
            cx.unique_indices.assign(p, p + np);
            p += np;
            if (mode == RIN_MODE_IA) {
                cx.unique_orientations.resize(np);
                for (uint32_t k = 0; k < np; ++k) cx.unique_orientations[k] = (*p++ != 0);
            }
        }
    }
    return true;
}

bool fetch_tet_maps(size_t n_tets, std::vector<long long>& global_vId_of_tet_vert,
    std::vector<size_t>& global_vId_start_index_of_tet, std::vector<size_t>& iso_fId_of_tet_face,
    std::vector<size_t>& iso_fId_start_index_of_tet, std::string& error)
{
    if (!g_ctx) {
        error = "no hot-path run";
        return false;
    }
    uint64_t na = 0, nv = 0, nf = 0;
    if (rin_tet_maps(g_ctx, &na, &nv, &nf) != RIN_OK) {
        error = rin_last_error();
        return false;
    }
    std::vector<uint32_t> act(na), voff(na + 1), foff(na + 1), fid(nf);
    std::vector<int64_t> vidv(nv);
    if (rin_download_tet_maps(g_ctx, act.data(), voff.data(), vidv.data(), foff.data(), fid.data()) != RIN_OK) {
        error = rin_last_error();
        return false;
    }
    global_vId_of_tet_vert.assign(vidv.begin(), vidv.end());
    iso_fId_of_tet_face.resize(nf);
    for (uint64_t i = 0; i < nf; ++i) iso_fId_of_tet_face[i] = widen(fid[i]);
    // inactive tets have empty ranges (src/extract_mesh.cpp:329-330); active tets are in tet order
    global_vId_start_index_of_tet.assign(n_tets + 1, 0);
    iso_fId_start_index_of_tet.assign(n_tets + 1, 0);
    uint64_t a = 0;
    for (size_t t = 0; t < n_tets; ++t) {
        const bool active = a < na && act[a] == t;
        global_vId_start_index_of_tet[t + 1] = global_vId_start_index_of_tet[t] + (active ? voff[a + 1] - voff[a] : 0);
        iso_fId_start_index_of_tet[t + 1] = iso_fId_start_index_of_tet[t] + (active ? foff[a + 1] - foff[a] : 0);
        if (active) ++a;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// N1: edges (device), patches and chains (host: order-dependent breadth-first traversals over a mesh that is tiny
// next to the tet grid)
// ---------------------------------------------------------------------------------------------------------------
bool mesh_edges(std::vector<std::vector<size_t>>& edges_of_face, std::vector<Edge>& edges, std::string& error)
{
    if (!ensure_ctx(error)) return false;
    uint64_t ne = 0;
    rin_counts n{};
    if (rin_mesh_edges(g_ctx, &ne) != RIN_OK || rin_get_counts(g_ctx, &n) != RIN_OK) {
        error = rin_last_error();
        return false;
    }
    std::vector<uint32_t> ev(2 * ne), eof(n.num_face_verts), eoff(ne + 1), ef(2 * n.num_face_verts), fo(n.num_faces + 1);
    rin_edges_out eo{ev.data(), eof.data(), eoff.data(), ef.data()};
    rin_mesh_out mo{};
    mo.face_offsets = fo.data();
    if (rin_download_edges(g_ctx, &eo) != RIN_OK || rin_download_mesh(g_ctx, &mo) != RIN_OK) {
        error = rin_last_error();
        return false;
    }
    edges.resize(ne);
    for (size_t e = 0; e < ne; ++e) {
        edges[e].v1 = ev[2 * e];
        edges[e].v2 = ev[2 * e + 1];
        edges[e].face_edge_indices.clear();
        edges[e].face_edge_indices.reserve(eoff[e + 1] - eoff[e]);
        for (uint32_t q = eoff[e]; q < eoff[e + 1]; ++q) edges[e].face_edge_indices.emplace_back(ef[2 * q], ef[2 * q + 1]);
    }
    edges_of_face.resize(n.num_faces);
    for (size_t f = 0; f < n.num_faces; ++f) edges_of_face[f].assign(eof.begin() + fo[f], eof.begin() + fo[f + 1]);
    return true;
}

namespace {
// breadth-first grouping of faces; `label_of(face)` is recorded for the seed of every patch
template <typename Label, typename LabelOf>
void group_patches(const std::vector<std::vector<size_t>>& edges_of_face, const std::vector<Edge>& edges,
    std::vector<std::vector<size_t>>& patches, std::vector<Label>& labels, LabelOf label_of)
{
    const size_t nf = edges_of_face.size();
    std::vector<char> seen(nf, 0);
    for (size_t seed = 0; seed < nf; ++seed) {
        if (seen[seed]) continue;
        patches.emplace_back();
        std::vector<size_t>& patch = patches.back(); // doubles as the queue: faces are appended in discovery order
        patch.push_back(seed);
        seen[seed] = 1;
        labels.push_back(label_of(seed));
        for (size_t head = 0; head < patch.size(); ++head) {
            const size_t f = patch[head];
            for (size_t e : edges_of_face[f]) {
                const auto& inc = edges[e].face_edge_indices;
                if (inc.size() != 2) continue; // only manifold edges connect the faces of a patch
                const size_t g = inc[0].first == f ? inc[1].first : inc[0].first;
                if (!seen[g]) {
                    seen[g] = 1;
                    patch.push_back(g);
                }
            }
        }
    }
}
} // namespace

void mesh_patches(const std::vector<std::vector<size_t>>& edges_of_face, const std::vector<Edge>& edges,
    const std::vector<PolygonFace>& faces, std::vector<std::vector<size_t>>& patches,
    std::vector<size_t>& patch_function_label)
{
    group_patches(edges_of_face, edges, patches, patch_function_label,
        [&](size_t f) { return faces[f].func_index.first; });
}

void mesh_patches(const std::vector<std::vector<size_t>>& edges_of_face, const std::vector<Edge>& edges,
    const std::vector<PolygonFace>& faces, std::vector<std::vector<size_t>>& patches,
    std::vector<std::pair<size_t, size_t>>& patch_function_label)
{
    group_patches(edges_of_face, edges, patches, patch_function_label, [&](size_t f) { return faces[f].func_index; });
}

void mesh_chains(size_t n_verts, const std::vector<Edge>& edges,
    std::vector<std::vector<size_t>>& non_manifold_edges_of_vert, std::vector<std::vector<size_t>>& chains)
{
    non_manifold_edges_of_vert.resize(n_verts);
    const size_t ne = edges.size();
    for (size_t e = 0; e < ne; ++e)
        if (edges[e].face_edge_indices.size() > 2) { // boundary edges (one face) take no part in the ordering
            non_manifold_edges_of_vert[edges[e].v1].push_back(e);
            non_manifold_edges_of_vert[edges[e].v2].push_back(e);
        }
    std::vector<char> seen(ne, 0);
    for (size_t seed = 0; seed < ne; ++seed) {
        if (seen[seed] || edges[seed].face_edge_indices.size() <= 2) continue;
        chains.emplace_back();
        std::vector<size_t>& chain = chains.back();
        chain.push_back(seed);
        seen[seed] = 1;
        for (size_t head = 0; head < chain.size(); ++head) {
            const size_t e = chain[head];
            for (const size_t v : {edges[e].v1, edges[e].v2}) {
                const auto& at = non_manifold_edges_of_vert[v];
                if (at.size() != 2) continue; // a junction (or a dangling end) closes the chain
                const size_t o = at[0] == e ? at[1] : at[0];
                if (!seen[o]) {
                    seen[o] = 1;
                    chain.push_back(o);
                }
            }
        }
    }
}

} // namespace rin_host
