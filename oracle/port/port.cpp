// ORACLE (test infrastructure, NOT product code; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library).
//
// CPU restatement ("port") of the reference's hot path, one function per reference stage:
//   orc_generate_grid    generate_tet_mesh          /root/reference/src/io.cpp:95-152
//   orc_eval_functions   load_functions             un-vendored duxingyi-charles/implicit_functions
//                                                   (cmake/implicit_functions.cmake:5-10), call site
//                                                   app/implicit_arrangement.cpp:57-64
//   orc_ia_run           implicit_arrangement       src/implicit_arrangement.cpp:53-402
//        signs :54-77, filter :86-116, per-tet dispatch :244-306,
//        extract_iso_mesh src/extract_mesh.cpp:10-265, compute_iso_vert_xyz :1446-1538
//   orc_mi_run           material_interface         src/material_interface.cpp:53-447
//        highest :59-92, filter :99-152, dispatch :288-351,
//        extract_MI_mesh src/extract_mesh.cpp:569-986, compute_MI_vert_xyz :1541-1637
// Per-tet engine: oracle/sa (restated simplicial_arrangement; PARITY UNPINNED at that boundary,
// pinned in aggregate by the reference's golden tests replayed through oracle/_ref).
// Single-threaded like the reference.  Compile with -ffp-contract=off.
#include <simplicial_arrangement/lookup_table.h>
#include <simplicial_arrangement/simplicial_arrangement.h>

#include "../../include/rin_b200.h"
#include "../result_bag.h"
#include "../sa/exact_arith.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <limits>
#include <map>
#include <cmath>
#include <cstring>
#include <set>
#include <unordered_map>

using namespace simplicial_arrangement;

namespace {

constexpr int64_t NONE64 = -1;

struct KeyHash
{
    size_t operator()(const std::array<uint64_t, 7>& a) const
    {
        uint64_t h = 1469598103934665603ULL;
        for (uint64_t x : a) {
            h ^= x;
            h *= 1099511628211ULL;
            h ^= h >> 29;
        }
        return size_t(h);
    }
};
using KeyMap = std::unordered_map<std::array<uint64_t, 7>, size_t, KeyHash>;

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// (smallest, second smallest, largest) of a list of distinct ids — src/extract_mesh.h:68-92
template <typename T>
std::array<T, 3> min2_max_key(const std::vector<T>& v)
{
    T mn = v[0], mx = v[0];
    size_t mn_pos = 0;
    for (size_t i = 1; i < v.size(); ++i) {
        if (v[i] < mn) {
            mn = v[i];
            mn_pos = i;
        } else if (v[i] > mx)
            mx = v[i];
    }
    T second = mx + 1;
    for (size_t i = 0; i < v.size(); ++i)
        if (i != mn_pos && v[i] < second) second = v[i];
    return {mn, second, mx};
}

double eval_one(const rin_func_desc& f, double x, double y, double z)
{
    double v = 0;
    switch (f.type) {
    case RIN_FN_PLANE: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        v = (f.p[3] * dx + f.p[4] * dy) + f.p[5] * dz;
        break;
    }
    case RIN_FN_SPHERE: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        double s = (dx * dx + dy * dy) + dz * dz;
        v = (f.p[4] != 0) ? f.p[3] * f.p[3] - s : f.p[3] - std::sqrt(s);
        break;
    }
    case RIN_FN_CYLINDER: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        double t = (f.p[3] * dx + f.p[4] * dy) + f.p[5] * dz;
        double px = dx - t * f.p[3], py = dy - t * f.p[4], pz = dz - t * f.p[5];
        v = f.p[6] - std::sqrt((px * px + py * py) + pz * pz);
        break;
    }
    case RIN_FN_TORUS: {
        double dx = x - f.p[0], dy = y - f.p[1], dz = z - f.p[2];
        double t = (f.p[3] * dx + f.p[4] * dy) + f.p[5] * dz;
        double px = dx - t * f.p[3], py = dy - t * f.p[4], pz = dz - t * f.p[5];
        double rho = std::sqrt((px * px + py * py) + pz * pz) - f.p[6];
        v = f.p[7] - std::sqrt(rho * rho + t * t);
        break;
    }
    default: v = 0; break;
    }
    return f.flip ? -v : v;
}

} // namespace

extern "C" {

// generate_tet_mesh — src/io.cpp:95-152 (vertex id = i*N*N + j*N + k, five tets per cube,
// two parities).  pts: N^3*3 doubles, tets: 5*R^3*4 uint64.
int orc_generate_grid(uint32_t R, const double* bmin, const double* bmax, double* pts, uint64_t* tets)
{
    if (R == 0) return -1;
    const uint64_t N = uint64_t(R) + 1;
    for (uint64_t i = 0; i < N; ++i) {
        double x = (double(i) / double(N - 1)) * (bmax[0] - bmin[0]) + bmin[0];
        for (uint64_t j = 0; j < N; ++j) {
            double y = (double(j) / double(N - 1)) * (bmax[1] - bmin[1]) + bmin[1];
            for (uint64_t k = 0; k < N; ++k) {
                double z = (double(k) / double(N - 1)) * (bmax[2] - bmin[2]) + bmin[2];
                double* p = pts + 3 * (i * N * N + j * N + k);
                p[0] = x;
                p[1] = y;
                p[2] = z;
            }
        }
    }
    // corner c of cube (i,j,k): bit0 -> +i, bit1 -> +j, bit2 -> +k, numbered as the reference's
    // v0..v7 (src/io.cpp:126-133)
    static const int di[8] = {0, 1, 1, 0, 0, 1, 1, 0};
    static const int dj[8] = {0, 0, 1, 1, 0, 0, 1, 1};
    static const int dk[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    static const int even[5][4] = {{4, 6, 1, 3}, {6, 3, 4, 7}, {1, 3, 0, 4}, {3, 1, 2, 6}, {4, 1, 6, 5}};
    static const int odd[5][4] = {{7, 0, 2, 5}, {2, 3, 0, 7}, {5, 7, 0, 4}, {7, 2, 6, 5}, {0, 1, 2, 5}};
    for (uint64_t i = 0; i < R; ++i)
        for (uint64_t j = 0; j < R; ++j)
            for (uint64_t k = 0; k < R; ++k) {
                uint64_t base = ((i * R + j) * R + k) * 5;
                const int(*tab)[4] = ((i + j + k) % 2 == 0) ? even : odd;
                for (int t = 0; t < 5; ++t)
                    for (int c = 0; c < 4; ++c) {
                        int q = tab[t][c];
                        tets[(base + t) * 4 + c] = (i + di[q]) * N * N + (j + dj[q]) * N + (k + dk[q]);
                    }
            }
    return 0;
}

// load_functions restated on flat descriptors; out is the reference's row-major V x F matrix
int orc_eval_functions(const rin_func_desc* funcs, uint32_t F, const double* pts, uint64_t V, double* out)
{
    for (uint64_t i = 0; i < V; ++i)
        for (uint32_t j = 0; j < F; ++j)
            out[i * F + j] = eval_one(funcs[j], pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    return 0;
}

// implicit_arrangement hot path.  flags: RIN_FLAG_*.  `tet_first/tet_count` restrict the run
// to a contiguous tet range (count 0 = all), mirroring the product's slab sharding.
void* orc_ia_run(const double* pts, uint64_t V, const uint64_t* tets_all, uint64_t T_all,
    const double* vals_in, uint32_t F, uint32_t flags, uint64_t tet_first, uint64_t tet_count)
{
    auto* bag = new ResultBag;
    bool use_lookup = flags & RIN_FLAG_USE_LOOKUP;
    bool use_secondary = (flags & RIN_FLAG_USE_SECONDARY_LOOKUP) && use_lookup; // :38-40
    if (use_lookup) {
        load_lookup_table(ARRANGEMENT);
        enable_lookup_table();
    } else
        disable_lookup_table(); // app/implicit_arrangement.cpp:39-42
    const uint64_t* tets = tets_all + 4 * tet_first;
    const uint64_t T = tet_count ? tet_count : T_all - tet_first;
    std::vector<double> negated;
    const double* vals = vals_in;
    if (flags & RIN_FLAG_NEGATE) { // src/csg.cpp:37
        negated.assign(vals_in, vals_in + V * F);
        for (double& x : negated) x = x * -1;
        vals = negated.data();
    }
    auto& tm = bag->f64["timings"];

    // ---- "func signs" :61-77
    double t0 = now_s();
    std::vector<int8_t> sgn(V * F);
    int64_t num_degenerate_vertex = 0;
    for (uint64_t i = 0; i < V * F; ++i) {
        double x = vals[i];
        sgn[i] = (x > 0) ? 1 : ((x < 0) ? -1 : 0);
        num_degenerate_vertex += (sgn[i] == 0);
    }
    tm.push_back(now_s() - t0);

    // ---- "filter" :86-116
    t0 = now_s();
    auto& func_in_tet = bag->i64["func_in_tet"];
    auto& start = bag->i64["start_index_of_tet"];
    start.reserve(T + 1);
    start.push_back(0);
    int64_t num_intersecting = 0;
    for (uint64_t t = 0; t < T; ++t) {
        const uint64_t* tv = tets + 4 * t;
        for (uint32_t j = 0; j < F; ++j) {
            int pos = 0, neg = 0;
            for (int c = 0; c < 4; ++c) {
                int s = sgn[tv[c] * F + j];
                pos += (s == 1);
                neg += (s == -1);
            }
            if (pos < 4 && neg < 4) func_in_tet.push_back(j);
        }
        if (int64_t(func_in_tet.size()) > start.back()) ++num_intersecting;
        start.push_back(int64_t(func_in_tet.size()));
    }
    tm.push_back(now_s() - t0);

    // ---- per-tet arrangements :244-306
    t0 = now_s();
    std::vector<Arrangement<3>> cuts;
    std::vector<int64_t> cut_index(T, NONE64);
    int64_t n1 = 0, n2 = 0, nmore = 0;
    try {
        std::vector<Plane<double, 3>> planes;
        for (uint64_t t = 0; t < T; ++t) {
            int64_t k = start[t + 1] - start[t];
            if (k == 0) continue;
            const uint64_t* tv = tets + 4 * t;
            planes.resize(k);
            for (int64_t j = 0; j < k; ++j) {
                int64_t f = func_in_tet[start[t] + j];
                for (int c = 0; c < 4; ++c) planes[j][c] = vals[tv[c] * F + f];
            }
            cut_index[t] = int64_t(cuts.size());
            if (use_lookup && !use_secondary && k == 2) {
                disable_lookup_table();
                cuts.emplace_back(compute_arrangement(planes));
                enable_lookup_table();
            } else
                cuts.emplace_back(compute_arrangement(planes));
            (k == 1 ? n1 : (k == 2 ? n2 : nmore))++;
        }
    } catch (std::runtime_error& e) {
        bag->error = e.what();
        return bag;
    }
    tm.push_back(now_s() - t0);

    // ---- "extract mesh": extract_iso_mesh, src/extract_mesh.cpp:10-265
    t0 = now_s();
    auto& vrec = bag->i64["vert_rec"]; // 10 per vertex: tet, local, size, sv[4], fi[3]
    auto& foff = bag->i64["face_offsets"];
    auto& fverts = bag->i64["face_verts"];
    auto& ftoff = bag->i64["face_tet_offsets"];
    auto& ftets = bag->i64["face_tets"];
    auto& ffunc = bag->i64["face_funcs"];
    std::vector<std::vector<int64_t>> face_vlist, face_tlist; // per face (faces may gain tets)
    KeyMap vert_map;
    std::unordered_map<std::array<uint64_t, 7>, size_t, KeyHash> face_map;
    size_t n_verts = 0;
    auto new_vert = [&](uint64_t tet, size_t local, int size, std::array<int64_t, 4> sv,
                        std::array<int64_t, 3> fi) {
        vrec.push_back(int64_t(tet));
        vrec.push_back(int64_t(local));
        vrec.push_back(size);
        for (auto x : sv) vrec.push_back(x);
        for (auto x : fi) vrec.push_back(x);
        return n_verts++;
    };
    std::vector<size_t> gid;
    std::vector<char> is_iso_v, is_iso_f;
    // maps of the cell-grouping overload (src/extract_mesh.cpp:268-566): same extraction, plus per tet the global
    // id of every local vertex (tet corners as -(id)-1, :402,518) and the iso-face id of every local face (:529-556)
    auto& tet_vmap = bag->i64["global_vId_of_tet_vert"];
    auto& tet_vstart = bag->i64["global_vId_start_index_of_tet"];
    auto& tet_fmap = bag->i64["iso_fId_of_tet_face"];
    auto& tet_fstart = bag->i64["iso_fId_start_index_of_tet"];
    tet_vstart.push_back(0);
    tet_fstart.push_back(0);
    for (uint64_t t = 0; t < T; ++t) {
        if (cut_index[t] == NONE64) {
            tet_vstart.push_back(int64_t(tet_vmap.size()));
            tet_fstart.push_back(int64_t(tet_fmap.size()));
            continue;
        }
        const auto& ar = cuts[cut_index[t]];
        const uint64_t* tv = tets + 4 * t;
        const int64_t s0 = start[t];
        // which faces / vertices are on an isosurface (:60-91)
        is_iso_v.assign(ar.vertices.size(), 0);
        is_iso_f.assign(ar.faces.size(), 0);
        for (size_t f = 0; f < ar.faces.size(); ++f) {
            bool iso = false;
            size_t sp = ar.faces[f].supporting_plane;
            if (ar.unique_planes.empty())
                iso = sp > 3;
            else
                for (size_t p : ar.unique_planes[ar.unique_plane_indices[sp]])
                    if (p > 3) {
                        iso = true;
                        break;
                    }
            if (iso) {
                is_iso_f[f] = 1;
                for (size_t v : ar.faces[f].vertices) is_iso_v[v] = 1;
            }
        }
        // iso-vertices (:93-230)
        gid.assign(ar.vertices.size(), size_t(-1));
        for (size_t j = 0; j < ar.vertices.size(); ++j) {
            std::array<int64_t, 3> fi = {NONE64, NONE64, NONE64};
            bool on_bndry[4] = {false, false, false, false};
            int nb = 0, ni = 0;
            for (size_t p : ar.vertices[j]) {
                if (p > 3)
                    fi[ni++] = func_in_tet[s0 + int64_t(p) - 4];
                else {
                    on_bndry[p] = true;
                    ++nb;
                }
            }
            if (!is_iso_v[j] || nb == 3) { // a tet corner: the one not on the three boundary planes (:389-402,:518)
                int corner = 0;
                while (on_bndry[corner]) ++corner;
                tet_vmap.push_back(-int64_t(tv[corner]) - 1);
            }
            if (!is_iso_v[j]) continue;
            // the minimal simplex containing the point = tet corners NOT on its boundary planes
            std::vector<uint64_t> corners;
            for (int c = 0; c < 4; ++c)
                if (!on_bndry[c]) corners.push_back(tv[c]);
            if (nb == 0) { // interior: never shared (:185-195), corners in tet order
                gid[j] = new_vert(t + tet_first, j, 4,
                    {int64_t(tv[0]), int64_t(tv[1]), int64_t(tv[2]), int64_t(tv[3])}, fi);
                tet_vmap.push_back(int64_t(gid[j]));
                continue;
            }
            std::sort(corners.begin(), corners.end());
            std::array<uint64_t, 7> key;
            key.fill(~0ULL);
            key[0] = corners.size();
            for (size_t c = 0; c < corners.size(); ++c) key[1 + c] = corners[c];
            // key function ids in plane-triple order, NOT sorted (:138,:167-168)
            for (int q = 0; q < ni; ++q) key[4 + q] = uint64_t(fi[q]);
            auto ins = vert_map.try_emplace(key, n_verts);
            if (ins.second) {
                std::array<int64_t, 4> sv = {NONE64, NONE64, NONE64, NONE64};
                for (size_t c = 0; c < corners.size(); ++c) sv[c] = int64_t(corners[c]);
                std::array<int64_t, 3> fstore = {NONE64, NONE64, NONE64};
                if (nb != 3) fstore = fi; // on-vertex case leaves func_indices untouched (:216-222)
                new_vert(t + tet_first, j, int(corners.size()), sv, fstore);
            }
            gid[j] = ins.first->second;
            if (nb != 3) tet_vmap.push_back(int64_t(gid[j]));
        }
        tet_vstart.push_back(int64_t(tet_vmap.size()));
        // iso-faces (:232-261)
        for (size_t f = 0; f < ar.faces.size(); ++f) {
            tet_fmap.push_back(NONE64);
            if (!is_iso_f[f]) continue;
            std::vector<int64_t> fv;
            for (size_t v : ar.faces[f].vertices) fv.push_back(int64_t(gid[v]));
            // QUIRK kept from the reference (:249,:258): for an iso-face coplanar with a tet
            // face the supporting plane is a boundary plane (< 4) and `supporting_plane - 4 +
            // start_index` wraps to an EARLIER entry of func_in_tet; out of range (UB there) -> None.
            int64_t fidx = s0 + int64_t(ar.faces[f].supporting_plane) - 4;
            int64_t fn = fidx >= 0 ? func_in_tet[fidx] : NONE64;
            if (ar.faces[f].negative_cell == Arrangement<3>::None) { // on the tet boundary
                auto k3 = min2_max_key(fv);
                std::array<uint64_t, 7> key;
                key.fill(~0ULL);
                key[0] = uint64_t(k3[0]);
                key[1] = uint64_t(k3[1]);
                key[2] = uint64_t(k3[2]);
                auto ins = face_map.try_emplace(key, face_vlist.size());
                if (!ins.second) {
                    face_tlist[ins.first->second].push_back(int64_t(t + tet_first));
                    face_tlist[ins.first->second].push_back(int64_t(f));
                    tet_fmap.back() = int64_t(ins.first->second);
                    continue;
                }
            }
            tet_fmap.back() = int64_t(face_vlist.size());
            face_vlist.push_back(fv);
            face_tlist.push_back({int64_t(t + tet_first), int64_t(f)});
            ffunc.push_back(fn);
            ffunc.push_back(NONE64);
        }
        tet_fstart.push_back(int64_t(tet_fmap.size()));
    }
    foff.push_back(0);
    ftoff.push_back(0);
    for (size_t f = 0; f < face_vlist.size(); ++f) {
        fverts.insert(fverts.end(), face_vlist[f].begin(), face_vlist[f].end());
        foff.push_back(int64_t(fverts.size()));
        ftets.insert(ftets.end(), face_tlist[f].begin(), face_tlist[f].end());
        ftoff.push_back(int64_t(ftets.size() / 2));
    }
    tm.push_back(now_s() - t0);

    // ---- "compute xyz": compute_iso_vert_xyz, src/extract_mesh.cpp:1446-1538
    t0 = now_s();
    auto& xyz = bag->f64["vert_xyz"];
    xyz.resize(3 * n_verts);
    auto P = [&](int64_t v, int c) { return pts[3 * v + c]; };
    auto Fv = [&](int64_t v, int64_t f) { return vals[uint64_t(v) * F + uint64_t(f)]; };
    for (size_t i = 0; i < n_verts; ++i) {
        const int64_t* r = &vrec[10 * i];
        const int64_t* sv = r + 3;
        const int64_t* fi = r + 7;
        double* o = &xyz[3 * i];
        switch (r[2]) {
        case 1:
            for (int c = 0; c < 3; ++c) o[c] = P(sv[0], c);
            break;
        case 2: { // src/extract_mesh.h:155-159
            double f1 = Fv(sv[0], fi[0]), f2 = Fv(sv[1], fi[0]);
            double b0 = f2 / (f2 - f1), b1 = 1 - b0;
            for (int c = 0; c < 3; ++c) o[c] = b0 * P(sv[0], c) + b1 * P(sv[1], c);
            break;
        }
        case 3: { // src/extract_mesh.h:139-151
            double a[3], b[3];
            for (int k = 0; k < 3; ++k) {
                a[k] = Fv(sv[k], fi[0]);
                b[k] = Fv(sv[k], fi[1]);
            }
            double n1 = a[2] * b[1] - a[1] * b[2];
            double n2 = a[0] * b[2] - a[2] * b[0];
            double n3 = a[1] * b[0] - a[0] * b[1];
            double d = n1 + n2 + n3;
            double w0 = n1 / d, w1 = n2 / d, w2 = n3 / d;
            for (int c = 0; c < 3; ++c) o[c] = w0 * P(sv[0], c) + w1 * P(sv[1], c) + w2 * P(sv[2], c);
            break;
        }
        case 4: { // src/extract_mesh.h:112-135
            double p1[4], p2[4], p3[4];
            for (int k = 0; k < 4; ++k) {
                p1[k] = Fv(sv[k], fi[0]);
                p2[k] = Fv(sv[k], fi[1]);
                p3[k] = Fv(sv[k], fi[2]);
            }
            double n1 = p1[3] * (p2[2] * p3[1] - p2[1] * p3[2]) + p1[2] * (p2[1] * p3[3] - p2[3] * p3[1]) +
                        p1[1] * (p2[3] * p3[2] - p2[2] * p3[3]);
            double n2 = p1[3] * (p2[0] * p3[2] - p2[2] * p3[0]) + p1[2] * (p2[3] * p3[0] - p2[0] * p3[3]) +
                        p1[0] * (p2[2] * p3[3] - p2[3] * p3[2]);
            double n3 = p1[3] * (p2[1] * p3[0] - p2[0] * p3[1]) + p1[1] * (p2[0] * p3[3] - p2[3] * p3[0]) +
                        p1[0] * (p2[3] * p3[1] - p2[1] * p3[3]);
            double n4 = p1[2] * (p2[0] * p3[1] - p2[1] * p3[0]) + p1[1] * (p2[2] * p3[0] - p2[0] * p3[2]) +
                        p1[0] * (p2[1] * p3[2] - p2[2] * p3[1]);
            double d = n1 + n2 + n3 + n4;
            double w[4] = {n1 / d, n2 / d, n3 / d, n4 / d};
            for (int c = 0; c < 3; ++c)
                o[c] = w[0] * P(sv[0], c) + w[1] * P(sv[1], c) + w[2] * P(sv[2], c) + w[3] * P(sv[3], c);
            break;
        }
        default: break;
        }
    }
    tm.push_back(now_s() - t0);

    auto es = engine_stats();
    bag->i64["stats"] = {int64_t(V), int64_t(T), num_degenerate_vertex, num_intersecting, n1, n2, nmore,
        int64_t(n_verts), int64_t(face_vlist.size())};
    bag->i64["engine"] = {int64_t(es.lookups), int64_t(es.general), int64_t(es.exact_fallbacks)};
    // full complexes are kept for orc_get_complex (tests of rin_get_complexes)
    bag->cuts_ia = std::move(cuts);
    bag->cut_index = std::move(cut_index);
    return bag;
}

// material_interface hot path (src/material_interface.cpp:53-447).
void* orc_mi_run(const double* pts, uint64_t V, const uint64_t* tets_all, uint64_t T_all,
    const double* vals, uint32_t F, uint32_t flags, uint64_t tet_first, uint64_t tet_count)
{
    auto* bag = new ResultBag;
    bool use_lookup = flags & RIN_FLAG_USE_LOOKUP;
    bool use_secondary = (flags & RIN_FLAG_USE_SECONDARY_LOOKUP) && use_lookup;
    if (use_lookup) {
        load_lookup_table(MATERIAL_INTERFACE);
        enable_lookup_table();
    } else
        disable_lookup_table(); // app/material_interface.cpp:42-45
    const uint64_t* tets = tets_all + 4 * tet_first;
    const uint64_t T = tet_count ? tet_count : T_all - tet_first;
    auto& tm = bag->f64["timings"];
    auto Fv = [&](uint64_t v, uint64_t f) { return vals[v * F + f]; };

    // ---- "highest func" :59-92
    double t0 = now_s();
    std::vector<uint32_t> highest(V);
    std::vector<char> degenerate(V, 0);
    std::unordered_map<uint64_t, std::vector<uint32_t>> highest_all;
    int64_t num_degenerate = 0;
    for (uint64_t i = 0; i < V; ++i) {
        double mx = Fv(i, 0);
        uint32_t id = 0, cnt = 1;
        for (uint32_t j = 1; j < F; ++j) {
            if (Fv(i, j) > mx) {
                mx = Fv(i, j);
                id = j;
                cnt = 1;
            } else if (Fv(i, j) == mx)
                ++cnt;
        }
        highest[i] = id;
        if (cnt > 1) {
            degenerate[i] = 1;
            ++num_degenerate;
            auto& l = highest_all[i];
            for (uint32_t j = 0; j < F; ++j)
                if (Fv(i, j) == mx) l.push_back(j);
        }
    }
    tm.push_back(now_s() - t0);

    // ---- "filter" :99-152
    t0 = now_s();
    auto& mat_in_tet = bag->i64["func_in_tet"];
    auto& start = bag->i64["start_index_of_tet"];
    start.push_back(0);
    int64_t num_intersecting = 0;
    {
        std::set<uint32_t> ms;
        for (uint64_t t = 0; t < T; ++t) {
            const uint64_t* tv = tets + 4 * t;
            ms.clear();
            for (int c = 0; c < 4; ++c) {
                if (degenerate[tv[c]]) {
                    const auto& l = highest_all[tv[c]];
                    ms.insert(l.begin(), l.end());
                } else
                    ms.insert(highest[tv[c]]);
            }
            if (ms.size() < 2) {
                start.push_back(int64_t(mat_in_tet.size()));
                continue;
            }
            double min_h[4];
            for (int c = 0; c < 4; ++c) {
                min_h[c] = std::numeric_limits<double>::max();
                for (uint32_t m : ms)
                    if (Fv(tv[c], m) < min_h[c]) min_h[c] = Fv(tv[c], m);
            }
            for (uint32_t j = 0; j < F; ++j) {
                int greater = 0;
                for (int c = 0; c < 4; ++c) greater += (Fv(tv[c], j) > min_h[c]);
                if (greater > 1) ms.insert(j);
            }
            ++num_intersecting;
            mat_in_tet.insert(mat_in_tet.end(), ms.begin(), ms.end());
            start.push_back(int64_t(mat_in_tet.size()));
        }
    }
    tm.push_back(now_s() - t0);

    // ---- per-tet material interfaces :288-351
    t0 = now_s();
    std::vector<MaterialInterface<3>> cuts;
    std::vector<int64_t> cut_index(T, NONE64);
    int64_t n2 = 0, n3 = 0, nmore = 0;
    try {
        std::vector<Material<double, 3>> mats;
        for (uint64_t t = 0; t < T; ++t) {
            int64_t k = start[t + 1] - start[t];
            if (k == 0) continue;
            const uint64_t* tv = tets + 4 * t;
            mats.resize(k);
            for (int64_t j = 0; j < k; ++j)
                for (int c = 0; c < 4; ++c) mats[j][c] = Fv(tv[c], mat_in_tet[start[t] + j]);
            cut_index[t] = int64_t(cuts.size());
            if (use_lookup && !use_secondary && k == 3) {
                disable_lookup_table();
                cuts.emplace_back(compute_material_interface(mats));
                enable_lookup_table();
            } else
                cuts.emplace_back(compute_material_interface(mats));
            (k == 2 ? n2 : (k == 3 ? n3 : nmore))++;
        }
    } catch (std::runtime_error& e) {
        bag->error = e.what();
        return bag;
    }
    tm.push_back(now_s() - t0);

    // ---- "extract mesh": extract_MI_mesh, src/extract_mesh.cpp:569-986
    t0 = now_s();
    auto& vrec = bag->i64["vert_rec"]; // 11 per vertex: tet, local, size, sv[4], mi[4]
    auto& ffunc = bag->i64["face_funcs"];
    std::vector<std::vector<int64_t>> face_vlist, face_tlist;
    KeyMap vert_map;
    size_t n_verts = 0;
    std::vector<int> vsize; // simplex size per MI vertex
    std::vector<int64_t> vsimplex0;
    auto new_vert = [&](uint64_t tet, size_t local, int size, std::array<int64_t, 4> sv, std::array<int64_t, 4> mi) {
        vrec.push_back(int64_t(tet));
        vrec.push_back(int64_t(local));
        vrec.push_back(size);
        for (auto x : sv) vrec.push_back(x);
        for (auto x : mi) vrec.push_back(x);
        vsize.push_back(size);
        vsimplex0.push_back(sv[0]);
        return n_verts++;
    };
    // boundary-face matching state (:603-605): face key -> material label(s) of the first side
    std::map<std::array<int64_t, 3>, std::vector<int64_t>> open_bfaces;
    std::map<std::array<int64_t, 3>, size_t> open_bface_slot; // ... and where its MI_fId_of_tet_face entry is
    std::vector<int64_t> lid;
    std::vector<char> is_mi_v, is_mi_f;
    // maps of the cell-grouping overload (src/extract_mesh.cpp:988-1443): per tet the global id of every local
    // vertex (tet corners as -(id)-1, :1130,1259) and the MI-face id of every local face; a boundary face that is
    // an interface between two tets gets the id in both (:1391-1392)
    auto& tet_vmap = bag->i64["global_vId_of_tet_vert"];
    auto& tet_vstart = bag->i64["global_vId_start_index_of_tet"];
    auto& tet_fmap = bag->i64["iso_fId_of_tet_face"];
    auto& tet_fstart = bag->i64["iso_fId_start_index_of_tet"];
    tet_vstart.push_back(0);
    tet_fstart.push_back(0);
    for (uint64_t t = 0; t < T; ++t) {
        if (cut_index[t] == NONE64) {
            tet_vstart.push_back(int64_t(tet_vmap.size()));
            tet_fstart.push_back(int64_t(tet_fmap.size()));
            continue;
        }
        const auto& mi = cuts[cut_index[t]];
        const uint64_t* tv = tets + 4 * t;
        const int64_t s0 = start[t];
        const uint64_t tg = t + tet_first;
        is_mi_v.assign(mi.vertices.size(), 0);
        is_mi_f.assign(mi.faces.size(), 0);
        for (size_t f = 0; f < mi.faces.size(); ++f)
            if (mi.faces[f].positive_material_label > 3) { // :641-651
                is_mi_f[f] = 1;
                for (size_t v : mi.faces[f].vertices) is_mi_v[v] = 1;
            }
        lid.assign(mi.vertices.size(), 0);
        for (size_t j = 0; j < mi.vertices.size(); ++j) {
            std::array<int64_t, 4> ms = {NONE64, NONE64, NONE64, NONE64};
            bool on_b[4] = {false, false, false, false};
            int nb = 0, nm = 0;
            for (size_t m : mi.vertices[j]) {
                if (m > 3)
                    ms[nm++] = mat_in_tet[s0 + int64_t(m) - 4];
                else {
                    on_b[m] = true;
                    ++nb;
                }
            }
            std::vector<uint64_t> corners;
            for (int c = 0; c < 4; ++c)
                if (!on_b[c]) corners.push_back(tv[c]);
            if (!is_mi_v[j]) { // a tet vertex that is not on the interface: encoded -(id)-1 (:670-687)
                lid[j] = -int64_t(corners[0]) - 1;
                continue;
            }
            if (nb == 0) { // :775-786
                lid[j] = int64_t(new_vert(tg, j, 4,
                    {int64_t(tv[0]), int64_t(tv[1]), int64_t(tv[2]), int64_t(tv[3])}, ms));
                continue;
            }
            std::sort(corners.begin(), corners.end());
            std::sort(ms.begin(), ms.begin() + nm); // material ids sorted in the key AND the record (:713-735,:751)
            std::array<uint64_t, 7> key;
            key.fill(~0ULL);
            key[0] = corners.size();
            for (size_t c = 0; c < corners.size(); ++c) key[1 + c] = corners[c];
            if (nb != 3)
                for (int q = 0; q < nm; ++q) key[4 + q] = uint64_t(ms[q]);
            auto ins = vert_map.try_emplace(key, n_verts);
            if (ins.second) {
                std::array<int64_t, 4> sv = {NONE64, NONE64, NONE64, NONE64};
                for (size_t c = 0; c < corners.size(); ++c) sv[c] = int64_t(corners[c]);
                std::array<int64_t, 4> mstore = {NONE64, NONE64, NONE64, NONE64};
                if (nb != 3) mstore = ms;
                new_vert(tg, j, int(corners.size()), sv, mstore);
            }
            lid[j] = int64_t(ins.first->second);
        }
        for (size_t j = 0; j < mi.vertices.size(); ++j)
            tet_vmap.push_back((lid[j] >= 0 && vsize[lid[j]] == 1) ? -vsimplex0[lid[j]] - 1 : lid[j]);
        tet_vstart.push_back(int64_t(tet_vmap.size()));
        for (size_t f = 0; f < mi.faces.size(); ++f) {
            const auto& face = mi.faces[f];
            tet_fmap.push_back(NONE64);
            if (is_mi_f[f]) { // :823-832
                std::vector<int64_t> fv;
                for (size_t v : face.vertices) fv.push_back(lid[v]);
                tet_fmap.back() = int64_t(face_vlist.size());
                face_vlist.push_back(fv);
                face_tlist.push_back({int64_t(tg), int64_t(f)});
                ffunc.push_back(mat_in_tet[s0 + int64_t(face.positive_material_label) - 4]);
                ffunc.push_back(mat_in_tet[s0 + int64_t(face.negative_material_label) - 4]);
                continue;
            }
            // simplex boundary face: match with the neighbouring tet (:833-950)
            std::vector<int64_t> bv;
            for (size_t v : face.vertices) {
                if (lid[v] >= 0 && vsize[lid[v]] == 1)
                    bv.push_back(-vsimplex0[lid[v]] - 1);
                else
                    bv.push_back(lid[v]);
            }
            auto key = min2_max_key(bv);
            std::vector<int64_t> labels;
            size_t nl = face.negative_material_label;
            if (mi.unique_materials.empty() || mi.unique_materials[mi.unique_material_indices[nl]].size() == 1)
                labels.push_back(mat_in_tet[s0 + int64_t(nl) - 4]);
            else
                for (size_t m : mi.unique_materials[mi.unique_material_indices[nl]])
                    labels.push_back(mat_in_tet[s0 + int64_t(m) - 4]);
            auto it = open_bfaces.find(key);
            if (it == open_bfaces.end()) {
                open_bfaces.emplace(key, labels);
                open_bface_slot[key] = tet_fmap.size() - 1;
                continue;
            }
            bool common = false;
            for (auto a : it->second)
                for (auto b : labels) common |= (a == b);
            if (common) {
                open_bfaces.erase(it); // same material on both sides: not an interface
                continue;
            }
            tet_fmap.back() = int64_t(face_vlist.size());
            tet_fmap[open_bface_slot[key]] = int64_t(face_vlist.size());
            // different materials on the two sides: the second tet emits the face (:952-981)
            std::vector<int64_t> fv;
            for (size_t k = 0; k < face.vertices.size(); ++k) {
                if (bv[k] < 0) {
                    uint64_t vid = uint64_t(-bv[k] - 1);
                    std::array<uint64_t, 7> vk;
                    vk.fill(~0ULL);
                    vk[0] = 1;
                    vk[1] = vid;
                    auto ins = vert_map.try_emplace(vk, n_verts);
                    if (ins.second)
                        new_vert(tg, face.vertices[k], 1, {int64_t(vid), NONE64, NONE64, NONE64},
                            {NONE64, NONE64, NONE64, NONE64});
                    fv.push_back(int64_t(ins.first->second));
                } else
                    fv.push_back(lid[face.vertices[k]]);
            }
            face_vlist.push_back(fv);
            face_tlist.push_back({int64_t(tg), int64_t(f)});
            // QUIRK kept from the reference (:979): positive label < 4 indexes BEFORE this tet's entries
            int64_t fidx = s0 + int64_t(face.positive_material_label) - 4;
            ffunc.push_back(fidx >= 0 ? mat_in_tet[fidx] : NONE64);
            ffunc.push_back(mat_in_tet[s0 + int64_t(nl) - 4]);
        }
        tet_fstart.push_back(int64_t(tet_fmap.size()));
    }
    auto& foff = bag->i64["face_offsets"];
    auto& fverts = bag->i64["face_verts"];
    auto& ftoff = bag->i64["face_tet_offsets"];
    auto& ftets = bag->i64["face_tets"];
    foff.push_back(0);
    ftoff.push_back(0);
    for (size_t f = 0; f < face_vlist.size(); ++f) {
        fverts.insert(fverts.end(), face_vlist[f].begin(), face_vlist[f].end());
        foff.push_back(int64_t(fverts.size()));
        ftets.insert(ftets.end(), face_tlist[f].begin(), face_tlist[f].end());
        ftoff.push_back(int64_t(ftets.size() / 2));
    }
    tm.push_back(now_s() - t0);

    // ---- "compute xyz": compute_MI_vert_xyz, src/extract_mesh.cpp:1541-1637
    t0 = now_s();
    auto& xyz = bag->f64["vert_xyz"];
    xyz.resize(3 * n_verts);
    auto P = [&](int64_t v, int c) { return pts[3 * v + c]; };
    for (size_t i = 0; i < n_verts; ++i) {
        const int64_t* r = &vrec[11 * i];
        const int64_t* sv = r + 3;
        const int64_t* m = r + 7;
        double* o = &xyz[3 * i];
        switch (r[2]) {
        case 1:
            for (int c = 0; c < 3; ++c) o[c] = P(sv[0], c);
            break;
        case 2: {
            double f1 = Fv(sv[0], m[0]) - Fv(sv[0], m[1]);
            double f2 = Fv(sv[1], m[0]) - Fv(sv[1], m[1]);
            double b0 = f2 / (f2 - f1), b1 = 1 - b0;
            for (int c = 0; c < 3; ++c) o[c] = b0 * P(sv[0], c) + b1 * P(sv[1], c);
            break;
        }
        case 3: {
            double a[3], b[3];
            for (int k = 0; k < 3; ++k) {
                a[k] = Fv(sv[k], m[0]) - Fv(sv[k], m[1]);
                b[k] = Fv(sv[k], m[1]) - Fv(sv[k], m[2]);
            }
            double n1 = a[2] * b[1] - a[1] * b[2];
            double n2_ = a[0] * b[2] - a[2] * b[0];
            double n3_ = a[1] * b[0] - a[0] * b[1];
            double d = n1 + n2_ + n3_;
            double w0 = n1 / d, w1 = n2_ / d, w2 = n3_ / d;
            for (int c = 0; c < 3; ++c) o[c] = w0 * P(sv[0], c) + w1 * P(sv[1], c) + w2 * P(sv[2], c);
            break;
        }
        case 4: {
            double p1[4], p2[4], p3[4];
            for (int k = 0; k < 4; ++k) {
                p1[k] = Fv(sv[k], m[0]) - Fv(sv[k], m[1]);
                p2[k] = Fv(sv[k], m[1]) - Fv(sv[k], m[2]);
                p3[k] = Fv(sv[k], m[2]) - Fv(sv[k], m[3]);
            }
            double n1 = p1[3] * (p2[2] * p3[1] - p2[1] * p3[2]) + p1[2] * (p2[1] * p3[3] - p2[3] * p3[1]) +
                        p1[1] * (p2[3] * p3[2] - p2[2] * p3[3]);
            double n2_ = p1[3] * (p2[0] * p3[2] - p2[2] * p3[0]) + p1[2] * (p2[3] * p3[0] - p2[0] * p3[3]) +
                         p1[0] * (p2[2] * p3[3] - p2[3] * p3[2]);
            double n3_ = p1[3] * (p2[1] * p3[0] - p2[0] * p3[1]) + p1[1] * (p2[0] * p3[3] - p2[3] * p3[0]) +
                         p1[0] * (p2[3] * p3[1] - p2[1] * p3[3]);
            double n4 = p1[2] * (p2[0] * p3[1] - p2[1] * p3[0]) + p1[1] * (p2[2] * p3[0] - p2[0] * p3[2]) +
                        p1[0] * (p2[1] * p3[2] - p2[2] * p3[1]);
            double d = n1 + n2_ + n3_ + n4;
            double w[4] = {n1 / d, n2_ / d, n3_ / d, n4 / d};
            for (int c = 0; c < 3; ++c)
                o[c] = w[0] * P(sv[0], c) + w[1] * P(sv[1], c) + w[2] * P(sv[2], c) + w[3] * P(sv[3], c);
            break;
        }
        default: break;
        }
    }
    tm.push_back(now_s() - t0);

    auto es = engine_stats();
    bag->i64["stats"] = {int64_t(V), int64_t(T), num_degenerate, num_intersecting, n2, n3, nmore, int64_t(n_verts),
        int64_t(face_vlist.size())};
    bag->i64["engine"] = {int64_t(es.lookups), int64_t(es.general), int64_t(es.exact_fallbacks)};
    bag->cuts_mi = std::move(cuts);
    bag->cut_index = std::move(cut_index);
    return bag;
}

const int64_t* orc_i64(void* h, const char* name, uint64_t* n)
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
const double* orc_f64(void* h, const char* name, uint64_t* n)
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
const char* orc_error(void* h)
{
    return static_cast<ResultBag*>(h)->error.c_str();
}
void orc_free(void* h)
{
    delete static_cast<ResultBag*>(h);
}

// serialise the complex of tet `t` of a finished run in the layout of rin_get_complexes
int orc_get_complex(void* h, uint64_t t, uint32_t* words, uint64_t cap, uint64_t* n_words)
{
    auto* b = static_cast<ResultBag*>(h);
    std::vector<uint32_t> w;
    if (t >= b->cut_index.size() || b->cut_index[t] < 0) {
        *n_words = 0;
        return 0;
    }
    auto none32 = [](size_t x) { return x == Arrangement<3>::None ? 0xffffffffu : uint32_t(x); };
    if (!b->cuts_ia.empty()) {
        const auto& a = b->cuts_ia[b->cut_index[t]];
        w = {uint32_t(a.vertices.size()), uint32_t(a.faces.size()), uint32_t(a.cells.size()),
            uint32_t(a.unique_planes.size())};
        for (auto& v : a.vertices)
            for (auto p : v) w.push_back(uint32_t(p));
        for (auto& f : a.faces) {
            w.push_back(uint32_t(f.supporting_plane));
            w.push_back(none32(f.positive_cell));
            w.push_back(none32(f.negative_cell));
            w.push_back(uint32_t(f.vertices.size()));
            for (auto v : f.vertices) w.push_back(uint32_t(v));
        }
        for (auto& c : a.cells) {
            w.push_back(uint32_t(c.faces.size()));
            for (auto f : c.faces) w.push_back(uint32_t(f));
        }
        if (!a.unique_planes.empty()) {
            w.push_back(uint32_t(a.unique_plane_indices.size()));
            for (auto g : a.unique_plane_indices) w.push_back(uint32_t(g));
            for (size_t p = 0; p < a.unique_plane_indices.size(); ++p)
                w.push_back(a.unique_plane_orientations[p] ? 1u : 0u);
        }
    } else {
        const auto& a = b->cuts_mi[b->cut_index[t]];
        w = {uint32_t(a.vertices.size()), uint32_t(a.faces.size()), uint32_t(a.cells.size()),
            uint32_t(a.unique_materials.size())};
        for (auto& v : a.vertices)
            for (auto p : v) w.push_back(uint32_t(p));
        for (auto& f : a.faces) {
            w.push_back(uint32_t(f.positive_material_label));
            w.push_back(uint32_t(f.negative_material_label));
            w.push_back(uint32_t(f.vertices.size()));
            for (auto v : f.vertices) w.push_back(uint32_t(v));
        }
        for (auto& c : a.cells) {
            w.push_back(uint32_t(c.material_label));
            w.push_back(uint32_t(c.faces.size()));
            for (auto f : c.faces) w.push_back(uint32_t(f));
        }
        if (!a.unique_materials.empty()) {
            w.push_back(uint32_t(a.unique_material_indices.size()));
            for (auto g : a.unique_material_indices) w.push_back(uint32_t(g));
        }
    }
    *n_words = w.size();
    if (words && cap >= w.size()) std::memcpy(words, w.data(), 4 * w.size());
    return 0;
}

// exact determinant sign (n <= 4, row-major) and the vertex-orientation predicate, for unit tests
int orc_det_sign(int n, const double* m)
{
    return sa_oracle::det_sign(n, m, nullptr);
}

// one per-tet arrangement, for unit tests of the engine: planes = k*4 doubles
int orc_compute_arrangement(const double* planes, uint32_t k, int use_lookup, uint32_t* words,
    uint64_t cap, uint64_t* n_words)
{
    ResultBag b;
    try {
        if (use_lookup) {
            load_lookup_table(ARRANGEMENT);
            enable_lookup_table();
        } else
            disable_lookup_table();
        std::vector<Plane<double, 3>> p(k);
        for (uint32_t j = 0; j < k; ++j)
            for (int c = 0; c < 4; ++c) p[j][c] = planes[4 * j + c];
        b.cuts_ia.push_back(compute_arrangement(p));
        b.cut_index = {0};
    } catch (std::runtime_error&) {
        return -5;
    }
    return orc_get_complex(&b, 0, words, cap, n_words);
}

int orc_compute_material_interface(const double* mats, uint32_t k, int use_lookup, uint32_t* words,
    uint64_t cap, uint64_t* n_words)
{
    ResultBag b;
    try {
        if (use_lookup) {
            load_lookup_table(MATERIAL_INTERFACE);
            enable_lookup_table();
        } else
            disable_lookup_table();
        std::vector<Material<double, 3>> m(k);
        for (uint32_t j = 0; j < k; ++j)
            for (int c = 0; c < 4; ++c) m[j][c] = mats[4 * j + c];
        b.cuts_mi.push_back(compute_material_interface(m));
        b.cut_index = {0};
    } catch (std::runtime_error&) {
        return -5;
    }
    return orc_get_complex(&b, 0, words, cap, n_words);
}

} // extern "C"
