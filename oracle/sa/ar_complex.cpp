// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of compute_arrangement() of qnzhou/simplicial_arrangement
// (un-vendored; /root/reference/cmake/simplicial_arrangement.cmake:5-10), called by the
// reference at /root/reference/src/implicit_arrangement.cpp:178,186,203,211,279,283.
// Algorithm (SIGGRAPH 2022 "Robust computation of implicit surface networks", sec. 4-5):
// start from the reference simplex (planes 0..3 are b_i = 0, positive inside), insert the
// input planes one at a time (plane id 4+j), classify every vertex of the current complex by
// the EXACT sign of the new plane at it, split crossed edges / faces / cells, record
// coincident planes.  Conventions (own; see DESIGN.md "per-tet complex conventions"):
//  * vertex plane triples are ascending; new vertices/edges/faces/cells are appended, entities
//    that were split are removed and the survivors keep their relative order;
//  * face loops are CCW seen from the positive side of their supporting plane, assuming the
//    simplex (v0,v1,v2,v3) is positively oriented;
//  * a new face's positive_cell / negative_cell lie on the positive / negative side of its plane;
//    simplex boundary faces have positive_cell = interior cell, negative_cell = None
//    (relied on at /root/reference/src/extract_mesh.cpp:240, src/topo_ray_shooting.cpp:587).
#include <simplicial_arrangement/lookup_table.h>
#include <simplicial_arrangement/simplicial_arrangement.h>

#include "ar_complex.h"
#include "exact_arith.h"

#include <algorithm>

namespace sa_oracle {

static simplicial_arrangement::EngineStats g_stats;
simplicial_arrangement::EngineStats& stats()
{
    return g_stats;
}

// exact sign of plane q at the intersection point of the three planes of vertex v
int ar_vertex_orientation(
    const std::vector<std::array<double, 4>>& planes, const std::array<size_t, 3>& v, const double* q)
{
    bool fixed[4] = {false, false, false, false};
    const double* impl[3];
    int ni = 0;
    for (int k = 0; k < 3; ++k) {
        if (v[k] < 4)
            fixed[v[k]] = true;
        else
            impl[ni++] = planes[v[k]].data();
    }
    int idx[4], n = 0;
    for (int c = 0; c < 4; ++c)
        if (!fixed[c]) idx[n++] = c;
    // n == ni + 1 : the point lives on an (n-1)-face of the simplex
    if (n == 1) {
        double x = q[idx[0]];
        return x > 0 ? 1 : (x < 0 ? -1 : 0);
    }
    double mq[16], md[16];
    for (int r = 0; r < ni; ++r)
        for (int c = 0; c < n; ++c) mq[r * n + c] = md[r * n + c] = impl[r][idx[c]];
    for (int c = 0; c < n; ++c) {
        mq[ni * n + c] = q[idx[c]];
        md[ni * n + c] = 1.0;
    }
    // q(X) = det[impl; q] / det[impl; 1]   (Cramer, sum of barycentric coordinates = 1)
    int sq = det_sign(n, mq, &g_stats.exact_fallbacks);
    if (sq == 0) return 0;
    int sd = det_sign(n, md, &g_stats.exact_fallbacks);
    if (sd == 0) throw std::runtime_error("simplicial_arrangement(oracle): degenerate vertex");
    return sq * sd;
}

void ARComplex::init_simplex()
{
    vertices = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
    edges.clear();
    // edge (a,b), a<b, lexicographic; supporting planes = the two other indices
    size_t eid[4][4];
    for (size_t a = 0; a < 4; ++a)
        for (size_t b = a + 1; b < 4; ++b) {
            size_t others[2], k = 0;
            for (size_t c = 0; c < 4; ++c)
                if (c != a && c != b) others[k++] = c;
            eid[a][b] = eid[b][a] = edges.size();
            edges.push_back({a, b, others[0], others[1]});
        }
    // face i: corners != i, CCW seen from inside (positive side of b_i = 0)
    static const size_t loops[4][3] = {{1, 3, 2}, {0, 2, 3}, {0, 3, 1}, {0, 1, 2}};
    faces.clear();
    for (size_t i = 0; i < 4; ++i) {
        Face f;
        for (int k = 0; k < 3; ++k) {
            f.verts.push_back(loops[i][k]);
            f.edges.push_back(eid[loops[i][k]][loops[i][(k + 1) % 3]]);
        }
        f.plane = i;
        f.pos_cell = 0;
        f.neg_cell = NONE;
        faces.push_back(f);
    }
    cells.assign(1, Cell{{0, 1, 2, 3}});
}

size_t ARComplex::add_plane(const std::vector<std::array<double, 4>>& planes, size_t pid)
{
    const double* q = planes[pid].data();
    // ---- step 1: vertices
    std::vector<int> o(vertices.size());
    bool any_nonzero = false;
    for (size_t i = 0; i < vertices.size(); ++i) {
        o[i] = ar_vertex_orientation(planes, vertices[i], q);
        any_nonzero |= (o[i] != 0);
    }
    if (!any_nonzero) throw std::runtime_error("simplicial_arrangement(oracle): null plane");

    // ---- step 2: edges
    struct ECut
    {
        size_t pos = NONE, neg = NONE, x = NONE;
        bool split = false;
    };
    const size_t nE = edges.size();
    std::vector<ECut> ecut(nE);
    for (size_t e = 0; e < nE; ++e) {
        const Edge E = edges[e];
        int o0 = o[E.v0], o1 = o[E.v1];
        ECut& c = ecut[e];
        if (o0 == 0 && o1 == 0) continue; // edge lies in the plane
        if (o0 == 0)
            c.x = E.v0;
        else if (o1 == 0)
            c.x = E.v1;
        if (o0 >= 0 && o1 >= 0)
            c.pos = e;
        else if (o0 <= 0 && o1 <= 0)
            c.neg = e;
        else {
            c.split = true;
            c.x = vertices.size();
            vertices.push_back({E.p0, E.p1, pid});
            o.push_back(0);
            Edge a{E.v0, c.x, E.p0, E.p1}, b{c.x, E.v1, E.p0, E.p1};
            c.pos = edges.size();
            c.neg = edges.size() + 1;
            if (o0 > 0) {
                edges.push_back(a);
                edges.push_back(b);
            } else {
                edges.push_back(b);
                edges.push_back(a);
            }
        }
    }

    // ---- step 3: faces
    struct FCut
    {
        size_t pos = NONE, neg = NONE, cut_edge = NONE;
        bool split = false;
    };
    const size_t nF = faces.size();
    std::vector<FCut> fcut(nF);
    size_t coplanar_plane = NONE;
    for (size_t f = 0; f < nF; ++f) {
        const Face F = faces[f];
        const size_t n = F.verts.size();
        size_t npos = 0, nneg = 0;
        for (size_t v : F.verts) {
            npos += (o[v] > 0);
            nneg += (o[v] < 0);
        }
        FCut& c = fcut[f];
        if (npos == 0 && nneg == 0) {
            if (coplanar_plane == NONE) coplanar_plane = F.plane;
            continue;
        }
        if (nneg == 0) {
            c.pos = f;
            continue;
        }
        if (npos == 0) {
            c.neg = f;
            continue;
        }
        c.split = true;
        auto O = [&](size_t k) { return o[F.verts[k % n]]; };
        size_t i = 0;
        while (!(O(i) <= 0 && O(i + 1) > 0)) ++i; // entering the positive run
        size_t jl = i + 1;
        while (O(jl + 1) > 0) ++jl; // last positive
        size_t start_tv, end_tv;
        Face P, N;
        P.plane = N.plane = F.plane;
        P.pos_cell = N.pos_cell = F.pos_cell;
        P.neg_cell = N.neg_cell = F.neg_cell;
        // positive loop
        const size_t ei = F.edges[i % n], ejl = F.edges[jl % n];
        size_t first_pos_edge, last_pos_edge, first_neg_edge, last_neg_edge;
        if (O(i) == 0) {
            start_tv = F.verts[i % n];
            first_pos_edge = ei;
            last_neg_edge = F.edges[(i + n - 1) % n];
        } else {
            start_tv = ecut[ei].x;
            first_pos_edge = ecut[ei].pos;
            last_neg_edge = ecut[ei].neg;
        }
        if (O(jl + 1) == 0) {
            end_tv = F.verts[(jl + 1) % n];
            last_pos_edge = ejl;
            first_neg_edge = F.edges[(jl + 1) % n];
        } else {
            end_tv = ecut[ejl].x;
            last_pos_edge = ecut[ejl].pos;
            first_neg_edge = ecut[ejl].neg;
        }
        const size_t ce = edges.size();
        edges.push_back({start_tv, end_tv, F.plane, pid});
        c.cut_edge = ce;
        P.verts.push_back(start_tv);
        P.edges.push_back(first_pos_edge);
        for (size_t k = i + 1; k <= jl; ++k) {
            P.verts.push_back(F.verts[k % n]);
            P.edges.push_back(k == jl ? last_pos_edge : F.edges[k % n]);
        }
        P.verts.push_back(end_tv);
        P.edges.push_back(ce);
        // negative loop: end_tv, negatives..., start_tv
        const size_t kfirst = (O(jl + 1) == 0) ? jl + 2 : jl + 1; // first strictly negative
        const size_t klast = (O(i) == 0) ? i + n - 1 : i + n; // last strictly negative
        N.verts.push_back(end_tv);
        N.edges.push_back(first_neg_edge);
        for (size_t k = kfirst; k <= klast; ++k) {
            N.verts.push_back(F.verts[k % n]);
            N.edges.push_back(k == klast ? last_neg_edge : F.edges[k % n]);
        }
        N.verts.push_back(start_tv);
        N.edges.push_back(ce);
        c.pos = faces.size();
        faces.push_back(P);
        c.neg = faces.size();
        faces.push_back(N);
    }

    // ---- step 4: cells
    const size_t nC = cells.size();
    std::vector<bool> cell_split(nC, false);
    for (size_t cidx = 0; cidx < nC; ++cidx) {
        const Cell C = cells[cidx];
        bool has_pos = false, has_neg = false;
        for (size_t f : C.faces) {
            has_pos |= (fcut[f].pos != NONE);
            has_neg |= (fcut[f].neg != NONE);
        }
        if (!(has_pos && has_neg)) continue;
        cell_split[cidx] = true;
        Cell CP, CN;
        // boundary edges of the cut polygon (+ which face/direction fixes the orientation)
        std::vector<size_t> cut_edges;
        size_t first_a = NONE, first_b = NONE; // first edge of the new face, direction a -> b
        auto add_cut_edge = [&](size_t e, size_t da, size_t db, bool inward, bool on_neg_side) {
            if (std::find(cut_edges.begin(), cut_edges.end(), e) != cut_edges.end()) return;
            cut_edges.push_back(e);
            if (first_a != NONE) return;
            // outward-CCW direction of e for this cell; the new face (whose own normal points
            // out of the negative sub-cell) runs opposite to it on the negative side.
            size_t oa = inward ? db : da, ob = inward ? da : db;
            if (on_neg_side) std::swap(oa, ob);
            first_a = oa;
            first_b = ob;
        };
        for (size_t f : C.faces) {
            const FCut& fc = fcut[f];
            const bool inward = (faces[f].pos_cell == cidx);
            if (fc.pos != NONE) CP.faces.push_back(fc.pos);
            if (fc.neg != NONE) CN.faces.push_back(fc.neg);
            if (fc.split) {
                const Edge& ce = edges[fc.cut_edge]; // runs start_tv -> end_tv in the neg loop
                add_cut_edge(fc.cut_edge, ce.v0, ce.v1, inward, true);
            } else if (fc.pos != NONE || fc.neg != NONE) {
                const Face& F = faces[f];
                const size_t n = F.verts.size();
                for (size_t k = 0; k < n; ++k) {
                    size_t a = F.verts[k], b = F.verts[(k + 1) % n];
                    if (o[a] == 0 && o[b] == 0)
                        add_cut_edge(F.edges[k], a, b, inward, fc.neg != NONE);
                }
            }
        }
        Face G;
        G.plane = pid;
        const size_t gid = faces.size();
        const size_t cp = cells.size(), cn = cells.size() + 1;
        G.pos_cell = cp;
        G.neg_cell = cn;
        // chain the edges into a loop starting with first_a -> first_b
        {
            std::vector<bool> used(cut_edges.size(), false);
            size_t cur = first_a;
            for (size_t step = 0; step < cut_edges.size(); ++step) {
                size_t pick = NONE;
                for (size_t k = 0; k < cut_edges.size(); ++k) {
                    if (used[k]) continue;
                    const Edge& E = edges[cut_edges[k]];
                    if (step == 0) {
                        if ((E.v0 == first_a && E.v1 == first_b) ||
                            (E.v1 == first_a && E.v0 == first_b)) {
                            pick = k;
                            break;
                        }
                    } else if (E.v0 == cur || E.v1 == cur) {
                        pick = k;
                        break;
                    }
                }
                if (pick == NONE)
                    throw std::runtime_error("simplicial_arrangement(oracle): open cut loop");
                used[pick] = true;
                const Edge& E = edges[cut_edges[pick]];
                G.verts.push_back(cur);
                G.edges.push_back(cut_edges[pick]);
                cur = (E.v0 == cur) ? E.v1 : E.v0;
            }
            if (cur != first_a)
                throw std::runtime_error("simplicial_arrangement(oracle): cut loop not closed");
        }
        faces.push_back(G);
        CP.faces.push_back(gid);
        CN.faces.push_back(gid);
        for (size_t f : CP.faces) {
            if (f == gid) continue;
            if (faces[f].pos_cell == cidx) faces[f].pos_cell = cp;
            if (faces[f].neg_cell == cidx) faces[f].neg_cell = cp;
        }
        for (size_t f : CN.faces) {
            if (f == gid) continue;
            if (faces[f].pos_cell == cidx) faces[f].pos_cell = cn;
            if (faces[f].neg_cell == cidx) faces[f].neg_cell = cn;
        }
        cells.push_back(CP);
        cells.push_back(CN);
    }

    // ---- step 5: consolidate (drop split entities, survivors keep their order)
    std::vector<size_t> emap(edges.size(), NONE), fmap(faces.size(), NONE), cmap(cells.size(), NONE);
    {
        size_t k = 0;
        for (size_t e = 0; e < edges.size(); ++e)
            if (!(e < nE && ecut[e].split)) emap[e] = k++;
        std::vector<Edge> ne(k);
        for (size_t e = 0; e < edges.size(); ++e)
            if (emap[e] != NONE) ne[emap[e]] = edges[e];
        edges.swap(ne);
        k = 0;
        for (size_t f = 0; f < faces.size(); ++f)
            if (!(f < nF && fcut[f].split)) fmap[f] = k++;
        size_t kc = 0;
        for (size_t c = 0; c < cells.size(); ++c)
            if (!(c < nC && cell_split[c])) cmap[c] = kc++;
        std::vector<Face> nf(k);
        for (size_t f = 0; f < faces.size(); ++f) {
            if (fmap[f] == NONE) continue;
            Face& F = faces[f];
            for (size_t& e : F.edges) e = emap[e];
            if (F.pos_cell != NONE) F.pos_cell = cmap[F.pos_cell];
            if (F.neg_cell != NONE) F.neg_cell = cmap[F.neg_cell];
            nf[fmap[f]] = std::move(F);
        }
        faces.swap(nf);
        std::vector<Cell> nc(kc);
        for (size_t c = 0; c < cells.size(); ++c) {
            if (cmap[c] == NONE) continue;
            for (size_t& f : cells[c].faces) f = fmap[f];
            nc[cmap[c]] = std::move(cells[c]);
        }
        cells.swap(nc);
    }
    return coplanar_plane;
}

simplicial_arrangement::Arrangement<3> compute_arrangement_general(
    const std::vector<std::array<double, 4>>& input)
{
    using simplicial_arrangement::Arrangement;
    std::vector<std::array<double, 4>> planes = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    planes.insert(planes.end(), input.begin(), input.end());
    ARComplex cx;
    cx.init_simplex();
    Arrangement<3> out;
    out.unique_plane_indices = {0, 1, 2, 3};
    out.unique_planes = {{0}, {1}, {2}, {3}};
    bool has_coplanar = false;
    for (size_t pid = 4; pid < planes.size(); ++pid) {
        size_t cop = cx.add_plane(planes, pid);
        if (cop == NONE) {
            out.unique_plane_indices.push_back(out.unique_planes.size());
            out.unique_planes.push_back({pid});
        } else {
            has_coplanar = true;
            size_t g = out.unique_plane_indices[cop];
            out.unique_plane_indices.push_back(g);
            out.unique_planes[g].push_back(pid);
        }
    }
    if (has_coplanar) {
        out.unique_plane_orientations.assign(planes.size(), true);
        for (size_t p = 0; p < planes.size(); ++p) {
            size_t r = out.unique_planes[out.unique_plane_indices[p]][0];
            if (r == p) continue;
            // p = c * r as linear functions; sign of c from any non-zero coefficient of r
            for (int k = 0; k < 4; ++k)
                if (planes[r][k] != 0) {
                    out.unique_plane_orientations[p] = ((planes[r][k] > 0) == (planes[p][k] > 0));
                    break;
                }
        }
    } else {
        out.unique_plane_indices.clear();
        out.unique_planes.clear();
    }
    out.vertices = cx.vertices;
    out.faces.resize(cx.faces.size());
    for (size_t f = 0; f < cx.faces.size(); ++f) {
        out.faces[f].vertices = cx.faces[f].verts;
        out.faces[f].supporting_plane = cx.faces[f].plane;
        out.faces[f].positive_cell = cx.faces[f].pos_cell;
        out.faces[f].negative_cell = cx.faces[f].neg_cell;
    }
    out.cells.resize(cx.cells.size());
    for (size_t c = 0; c < cx.cells.size(); ++c) out.cells[c].faces = cx.cells[c].faces;
    return out;
}

} // namespace sa_oracle
