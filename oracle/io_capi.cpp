// ORACLE (test infrastructure): one synthetic result set written through either the reference's own writers
// (IO_REF: /root/reference/src/io.cpp compiled in place) or the product's (host/rin_io.cpp), for the byte-for-byte
// comparison in tests/test_io_writers.py.  Also dumps what parse_config_file / load_tet_mesh read.
#ifdef IO_REF
#include "io.h"
#define NS
#else
#include "../robust-implicit-surface-networks_b200/host/rin_io.h"
#define NS rin_host::
using rin_host::Config;
#endif

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

namespace {
uint64_t sm64(uint64_t& s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
double number(uint64_t& s)
{
    // coordinates as they occur (uniform in [-1, 1]), plus integers, tiny, huge and awkward values
    const uint64_t r = sm64(s);
    const double u = (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0);
    switch (r % 16) {
    case 0: return double(int64_t(sm64(s) % 2001) - 1000);
    case 1: return std::ldexp(u, int(sm64(s) % 120) - 60);
    case 2: return -std::ldexp(u, int(sm64(s) % 600) - 300);
    case 3: return (sm64(s) % 3 == 0) ? 0.0 : (sm64(s) % 2 ? -0.0 : 1e15 * u);
    case 4: return double(sm64(s) % 100000) * 1e-4;
    default: return 2 * u - 1;
    }
}
} // namespace

extern "C" int io_write_all(const char* dir, uint64_t seed, int n_pts, int n_faces)
{
    uint64_t s = seed;
    std::vector<std::array<double, 3>> pts(n_pts);
    for (auto& p : pts)
        for (auto& x : p) x = number(s);
    std::vector<PolygonFace> faces(n_faces);
    for (auto& f : faces) {
        const int k = 3 + int(sm64(s) % 4);
        for (int j = 0; j < k; ++j) f.vert_indices.push_back(n_pts ? sm64(s) % n_pts : 0);
        f.func_index = {sm64(s) % 8, sm64(s) % 8};
    }
    std::vector<std::vector<size_t>> patches, chains, nme(n_pts), shells, cells;
    for (int f = 0; f < n_faces;) {
        patches.emplace_back();
        const int len = 1 + int(sm64(s) % 7);
        for (int j = 0; j < len && f < n_faces; ++j) patches.back().push_back(f++);
    }
    std::vector<Edge> edges(n_pts ? n_pts / 2 + 1 : 0);
    for (auto& e : edges) {
        e.v1 = sm64(s) % n_pts;
        e.v2 = sm64(s) % n_pts;
    }
    for (size_t e = 0; e < edges.size();) {
        chains.emplace_back();
        const int len = 1 + int(sm64(s) % 5);
        for (int j = 0; j < len && e < edges.size(); ++j) chains.back().push_back(e++);
        if (sm64(s) % 3 == 0) e += 2; // not every edge is in a chain
    }
    for (auto& l : nme) {
        const int k = int(sm64(s) % 5);
        for (int j = 0; j < k; ++j) l.push_back(sm64(s) % (edges.size() + 1));
    }
    for (size_t p = 0; p < 2 * patches.size();) {
        shells.emplace_back();
        const int len = 1 + int(sm64(s) % 4);
        for (int j = 0; j < len && p < 2 * patches.size(); ++j) shells.back().push_back(p++);
    }
    for (size_t sh = 0; sh < shells.size();) {
        cells.emplace_back();
        const int len = 1 + int(sm64(s) % 2);
        for (int j = 0; j < len && sh < shells.size(); ++j) cells.back().push_back(sh++);
    }
    std::vector<size_t> plabel_ia, clabel_mi;
    std::vector<std::pair<size_t, size_t>> plabel_mi;
    std::vector<bool> psign;
    for (size_t p = 0; p < patches.size(); ++p) {
        plabel_ia.push_back(sm64(s) % 8);
        plabel_mi.emplace_back(sm64(s) % 8, sm64(s) % 3 ? sm64(s) % 8 : ~size_t(0));
        psign.push_back(sm64(s) % 2);
    }
    std::vector<std::vector<bool>> clabel_ia;
    for (size_t c = 0; c < cells.size(); ++c) {
        clabel_ia.emplace_back();
        for (int f = 0; f < 5; ++f) clabel_ia.back().push_back(sm64(s) % 2);
        clabel_mi.push_back(sm64(s) % 8);
    }
    const std::string d(dir);
    bool ok = NS save_result(d + "/mesh.json", pts, faces, patches, plabel_ia, edges, chains, nme, shells, cells, clabel_ia);
    ok &= NS save_result_MI(d + "/mesh_mi.json", pts, faces, patches, plabel_mi, edges, chains, nme, shells, cells, clabel_mi);
    ok &= NS save_result_CSG(d + "/mesh_csg.json", pts, faces, patches, psign, edges, chains, nme);
    ok &= NS save_result_msh(d + "/mesh", pts, faces, patches, edges, chains, nme, shells, cells);
    std::vector<std::string> tl = {"func signs", "filter", "simp_arr(other)", "extract mesh", "compute xyz", "patches",
        "arrCells(other)", "filter"};
    std::vector<double> tv;
    for (size_t i = 0; i < tl.size(); ++i) tv.push_back(std::fabs(number(s)) * 1e-3);
    std::vector<std::string> sl = {"num_pts", "num_tets", "num_iso_verts", "num_cells", "num_1_func"};
    std::vector<size_t> sv;
    for (size_t i = 0; i < sl.size(); ++i) sv.push_back(sm64(s) % 100000000);
    if (n_pts == 0) {
        tl.clear();
        tv.clear();
        sl.clear();
        sv.clear();
    }
    ok &= NS save_timings(d + "/timings.json", tl, tv);
    ok &= NS save_statistics(d + "/stats.json", sl, sv);
    return ok ? 0 : 1;
}

// parse_config_file + load_tet_mesh: everything they read, as text
extern "C" int io_read_all(const char* config_file, char* out, int cap)
{
    std::ostringstream o;
    o.precision(17);
    try {
        const Config c = NS parse_config_file(config_file);
        o << c.tet_mesh_file << "|" << c.func_file << "|" << c.output_dir << "|" << c.use_lookup << c.use_secondary_lookup
          << c.use_topo_ray_shooting << "|" << c.tet_mesh_resolution;
        if (c.tet_mesh_file.empty())
            for (int k = 0; k < 3; ++k) o << "|" << c.tet_mesh_bbox_min[k] << "," << c.tet_mesh_bbox_max[k];
        else {
            std::vector<std::array<double, 3>> pts;
            std::vector<std::array<size_t, 4>> tets;
            const bool ok = NS load_tet_mesh(c.tet_mesh_file, pts, tets);
            o << "|mesh " << ok << " " << pts.size() << " " << tets.size();
            for (auto& p : pts) o << " " << p[0] << " " << p[1] << " " << p[2];
            for (auto& t : tets) o << " " << t[0] << " " << t[1] << " " << t[2] << " " << t[3];
        }
    } catch (const std::exception& e) {
        o << "exception: " << e.what();
    }
    const std::string s = o.str();
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
