// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement of compute_material_interface() of qnzhou/simplicial_arrangement
// (un-vendored; /root/reference/cmake/simplicial_arrangement.cmake:5-10), called by the reference
// at /root/reference/src/material_interface.cpp:218-327.
// Algorithm (SIGGRAPH 2022 paper, sec. 5.2): materials are inserted one at a time (material id
// 4+j; ids 0..3 are pseudo materials "outside simplex face i"); every cell carries the material
// that is maximal in it.  Inserting M classifies every vertex by the EXACT sign of (M - current
// maximum) there, cuts every cell by the plane M = its material, and merges all positive parts
// into the single (convex) cell of M, dropping everything strictly inside it.
// Conventions (own; DESIGN.md "per-tet complex conventions"):
//  * vertex material quadruples ascend; a face (positive_material_label, negative_material_label)
//    is a closed loop CCW seen from the positive material's side; simplex boundary faces have
//    positive label = face id (<= 3) and negative label = the cell inside
//    (relied on at /root/reference/src/extract_mesh.cpp:643,850, src/topo_ray_shooting.cpp:769,985);
//    interface faces have positive label > negative label (later material on the positive side);
//  * survivors keep their order; new entities are appended: split parts, cut edges / cut faces,
//    merged simplex-edge pieces, merged boundary faces, negative sub-cells, then the new cell.
#include <simplicial_arrangement/lookup_table.h>
#include <simplicial_arrangement/simplicial_arrangement.h>

#include "ar_complex.h"
#include "exact_arith.h"

#include <algorithm>
#include <map>

namespace sa_oracle {

namespace {

struct MIComplex
{
    struct Edge
    {
        size_t v0, v1;
        std::array<size_t, 3> m; // materials equal along the edge (ascending)
    };
    struct Face
    {
        std::vector<size_t> verts, edges; // loop; edges[k] joins verts[k], verts[k+1]
        size_t bface = NONE;              // simplex face id if the face lies on the boundary
        size_t pos_cell = NONE, neg_cell = NONE; // pos_cell == NONE on the boundary
    };
    struct Cell
    {
        std::vector<size_t> faces;
        size_t material = NONE;
    };
    std::vector<std::array<size_t, 4>> vertices;
    std::vector<Edge> edges;
    std::vector<Face> faces;
    std::vector<Cell> cells;

    void init();
    size_t add_material(const std::vector<std::array<double, 4>>& mats, size_t mid);
};

// exact sign of (M - current maximum) at vertex v (the point where its real materials tie)
int mi_vertex_orientation(
    const std::vector<std::array<double, 4>>& mats, const std::array<size_t, 4>& v, const double* M)
{
    bool fixed[4] = {false, false, false, false};
    const double* real[4];
    int k = 0;
    for (size_t id : v) {
        if (id < 4)
            fixed[id] = true;
        else
            real[k++] = mats[id].data();
    }
    int idx[4], n = 0;
    for (int c = 0; c < 4; ++c)
        if (!fixed[c]) idx[n++] = c;
    // n == k : the point lives on an (n-1)-face of the simplex, k real materials tie there
    if (n == 1) {
        double a = M[idx[0]], b = real[0][idx[0]];
        return a > b ? 1 : (a < b ? -1 : 0);
    }
    // rows r_i - r_{i+1} (i < k-1) and the query row M - r_0 (or all ones): entries a - b
    double qa[16], qb[16], da[16], db[16];
    for (int r = 0; r < k - 1; ++r)
        for (int c = 0; c < n; ++c) {
            qa[r * n + c] = da[r * n + c] = real[r][idx[c]];
            qb[r * n + c] = db[r * n + c] = real[r + 1][idx[c]];
        }
    for (int c = 0; c < n; ++c) {
        qa[(k - 1) * n + c] = M[idx[c]];
        qb[(k - 1) * n + c] = real[0][idx[c]];
        da[(k - 1) * n + c] = 1.0;
        db[(k - 1) * n + c] = 0.0;
    }
    int sq = det_sign_diff(n, qa, qb, &stats().exact_fallbacks);
    if (sq == 0) return 0;
    int sd = det_sign_diff(n, da, db, &stats().exact_fallbacks);
    if (sd == 0) throw std::runtime_error("simplicial_arrangement(oracle): degenerate MI vertex");
    return sq * sd;
}

void MIComplex::init()
{
    vertices = {{1, 2, 3, 4}, {0, 2, 3, 4}, {0, 1, 3, 4}, {0, 1, 2, 4}};
    edges.clear();
    size_t eid[4][4];
    for (size_t a = 0; a < 4; ++a)
        for (size_t b = a + 1; b < 4; ++b) {
            size_t others[2], k = 0;
            for (size_t c = 0; c < 4; ++c)
                if (c != a && c != b) others[k++] = c;
            eid[a][b] = eid[b][a] = edges.size();
            edges.push_back({a, b, {others[0], others[1], 4}});
        }
    // boundary face i: CCW seen from outside (the positive, pseudo-material side)
    static const size_t loops[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
    faces.clear();
    for (size_t i = 0; i < 4; ++i) {
        Face f;
        for (int k = 0; k < 3; ++k) {
            f.verts.push_back(loops[i][k]);
            f.edges.push_back(eid[loops[i][k]][loops[i][(k + 1) % 3]]);
        }
        f.bface = i;
        f.pos_cell = NONE;
        f.neg_cell = 0;
        faces.push_back(f);
    }
    cells.assign(1, Cell{{0, 1, 2, 3}, 4});
}

// returns the material M coincides with (identical on a whole cell), or NONE
size_t MIComplex::add_material(const std::vector<std::array<double, 4>>& mats, size_t mid)
{
    const double* M = mats[mid].data();
    std::vector<int> o(vertices.size());
    for (size_t i = 0; i < vertices.size(); ++i) o[i] = mi_vertex_orientation(mats, vertices[i], M);

    // ---- edges
    struct ECut
    {
        size_t pos = NONE, neg = NONE, x = NONE;
        bool split = false;
    };
    const size_t nV0 = vertices.size();
    const size_t nE = edges.size();
    std::vector<ECut> ecut(nE);
    for (size_t e = 0; e < nE; ++e) {
        const Edge E = edges[e];
        int o0 = o[E.v0], o1 = o[E.v1];
        ECut& c = ecut[e];
        if (o0 == 0 && o1 == 0) continue;
        if (o0 >= 0 && o1 >= 0)
            c.pos = e;
        else if (o0 <= 0 && o1 <= 0)
            c.neg = e;
        else {
            c.split = true;
            c.x = vertices.size();
            vertices.push_back({E.m[0], E.m[1], E.m[2], mid});
            o.push_back(0);
            Edge a{E.v0, c.x, E.m}, b{c.x, E.v1, E.m};
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
    // ---- faces
    struct FCut
    {
        size_t pos = NONE, neg = NONE, cut_edge = NONE;
        bool split = false;
    };
    const size_t nF = faces.size();
    std::vector<FCut> fcut(nF);
    auto face_materials = [&](const Face& F) { // the two materials meeting at F (ascending)
        size_t a = (F.pos_cell == NONE) ? F.bface : cells[F.pos_cell].material;
        size_t b = cells[F.neg_cell].material;
        return std::array<size_t, 2>{std::min(a, b), std::max(a, b)};
    };
    for (size_t f = 0; f < nF; ++f) {
        const Face F = faces[f];
        const size_t n = F.verts.size();
        size_t npos = 0, nneg = 0;
        for (size_t v : F.verts) {
            npos += (o[v] > 0);
            nneg += (o[v] < 0);
        }
        FCut& c = fcut[f];
        if (npos == 0 && nneg == 0) continue; // lies in M = current
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
        while (!(O(i) <= 0 && O(i + 1) > 0)) ++i;
        size_t jl = i + 1;
        while (O(jl + 1) > 0) ++jl;
        size_t start_tv, end_tv, first_pos, last_pos, first_neg, last_neg;
        const size_t ei = F.edges[i % n], ejl = F.edges[jl % n];
        if (O(i) == 0) {
            start_tv = F.verts[i % n];
            first_pos = ei;
            last_neg = F.edges[(i + n - 1) % n];
        } else {
            start_tv = ecut[ei].x;
            first_pos = ecut[ei].pos;
            last_neg = ecut[ei].neg;
        }
        if (O(jl + 1) == 0) {
            end_tv = F.verts[(jl + 1) % n];
            last_pos = ejl;
            first_neg = F.edges[(jl + 1) % n];
        } else {
            end_tv = ecut[ejl].x;
            last_pos = ecut[ejl].pos;
            first_neg = ecut[ejl].neg;
        }
        const size_t ce = edges.size();
        auto fm = face_materials(F);
        edges.push_back({start_tv, end_tv, {fm[0], fm[1], mid}});
        c.cut_edge = ce;
        Face P, N;
        P.bface = N.bface = F.bface;
        P.pos_cell = N.pos_cell = F.pos_cell;
        P.neg_cell = N.neg_cell = F.neg_cell;
        P.verts.push_back(start_tv);
        P.edges.push_back(first_pos);
        for (size_t k = i + 1; k <= jl; ++k) {
            P.verts.push_back(F.verts[k % n]);
            P.edges.push_back(k == jl ? last_pos : F.edges[k % n]);
        }
        P.verts.push_back(end_tv);
        P.edges.push_back(ce);
        const size_t kfirst = (O(jl + 1) == 0) ? jl + 2 : jl + 1;
        const size_t klast = (O(i) == 0) ? i + n - 1 : i + n;
        N.verts.push_back(end_tv);
        N.edges.push_back(first_neg);
        for (size_t k = kfirst; k <= klast; ++k) {
            N.verts.push_back(F.verts[k % n]);
            N.edges.push_back(k == klast ? last_neg : F.edges[k % n]);
        }
        N.verts.push_back(start_tv);
        N.edges.push_back(ce);
        c.pos = faces.size();
        faces.push_back(P);
        c.neg = faces.size();
        faces.push_back(N);
    }
    // ---- cells: positive parts (to be merged), negative parts, duplicates
    const size_t nC = cells.size();
    enum { C_NEG = 0, C_POS = 1, C_SPLIT = 2, C_ZERO = 3 };
    std::vector<int> cstat(nC, C_NEG);
    std::vector<size_t> pos_part_faces;            // faces of all positive parts (with repeats)
    std::vector<std::pair<size_t, bool>> pos_side; // (face, the positive part is on its pos side)
    std::vector<Cell> neg_subcells;
    std::vector<size_t> neg_subcell_of(nC, NONE), cut_face_of(nC, NONE);
    size_t duplicate_of = NONE;
    const size_t first_cut_face = faces.size();
    for (size_t cidx = 0; cidx < nC; ++cidx) {
        const Cell C = cells[cidx];
        bool has_pos = false, has_neg = false;
        for (size_t f : C.faces) {
            has_pos |= (fcut[f].pos != NONE);
            has_neg |= (fcut[f].neg != NONE);
        }
        if (!has_pos && !has_neg) {
            cstat[cidx] = C_ZERO;
            if (duplicate_of == NONE) duplicate_of = C.material;
            continue;
        }
        if (!has_pos) continue;
        if (!has_neg) {
            cstat[cidx] = C_POS;
            continue;
        }
        cstat[cidx] = C_SPLIT;
        Cell CN;
        CN.material = C.material;
        std::vector<size_t> cut_edges;
        size_t first_a = NONE, first_b = NONE;
        auto add_cut_edge = [&](size_t e, size_t da, size_t db, bool inward, bool on_neg_side) {
            if (std::find(cut_edges.begin(), cut_edges.end(), e) != cut_edges.end()) return;
            cut_edges.push_back(e);
            if (first_a != NONE) return;
            size_t oa = inward ? db : da, ob = inward ? da : db;
            if (on_neg_side) std::swap(oa, ob);
            first_a = oa;
            first_b = ob;
        };
        for (size_t f : C.faces) {
            const FCut& fc = fcut[f];
            const bool inward = (faces[f].pos_cell == cidx);
            if (fc.neg != NONE) CN.faces.push_back(fc.neg);
            if (fc.split) {
                const Edge& ce = edges[fc.cut_edge];
                add_cut_edge(fc.cut_edge, ce.v0, ce.v1, inward, true);
            } else if (fc.pos != NONE || fc.neg != NONE) {
                const Face& F = faces[f];
                const size_t n = F.verts.size();
                for (size_t k = 0; k < n; ++k) {
                    size_t a = F.verts[k], b = F.verts[(k + 1) % n];
                    if (o[a] == 0 && o[b] == 0) add_cut_edge(F.edges[k], a, b, inward, fc.neg != NONE);
                }
            }
        }
        Face G; // pos side: the new material's cell (filled in below), neg side: the negative part
        {
            std::vector<bool> used(cut_edges.size(), false);
            size_t cur = first_a;
            for (size_t step = 0; step < cut_edges.size(); ++step) {
                size_t pick = NONE;
                for (size_t k = 0; k < cut_edges.size(); ++k) {
                    if (used[k]) continue;
                    const Edge& E = edges[cut_edges[k]];
                    if (step == 0) {
                        if ((E.v0 == first_a && E.v1 == first_b) || (E.v1 == first_a && E.v0 == first_b)) {
                            pick = k;
                            break;
                        }
                    } else if (E.v0 == cur || E.v1 == cur) {
                        pick = k;
                        break;
                    }
                }
                if (pick == NONE) throw std::runtime_error("simplicial_arrangement(oracle): open MI cut loop");
                used[pick] = true;
                const Edge& E = edges[cut_edges[pick]];
                G.verts.push_back(cur);
                G.edges.push_back(cut_edges[pick]);
                cur = (E.v0 == cur) ? E.v1 : E.v0;
            }
            if (cur != first_a) throw std::runtime_error("simplicial_arrangement(oracle): MI cut loop not closed");
        }
        cut_face_of[cidx] = faces.size();
        faces.push_back(G);
        CN.faces.push_back(cut_face_of[cidx]);
        neg_subcell_of[cidx] = neg_subcells.size();
        neg_subcells.push_back(CN);
    }
    const size_t n_faces_after_cut = faces.size();
    (void)first_cut_face;
    bool any_pos = false;
    for (int s : cstat) any_pos |= (s == C_POS || s == C_SPLIT);

    // ---- rebuild: drop what lies strictly inside the new material's region, merge its boundary
    // status of the entities that exist now
    auto edge_is_positive = [&](size_t e) { // lies in the closed positive region, not in M = current
        if (e < nE) return !ecut[e].split && ecut[e].pos == e;
        const Edge& E = edges[e];
        return (o[E.v0] > 0 || o[E.v1] > 0);
    };
    auto on_simplex_edge = [&](const Edge& E) { return E.m[1] < 4; }; // two boundary materials
    auto face_is_positive = [&](size_t f) {
        if (f < nF) return !fcut[f].split && fcut[f].pos == f;
        if (f >= n_faces_after_cut) return false;
        for (size_t v : faces[f].verts)
            if (o[v] > 0) return true;
        return false;
    };
    // vertices: strictly positive ones disappear unless they are simplex corners
    std::vector<size_t> vmap(vertices.size(), NONE);
    {
        size_t k = 0;
        for (size_t v = 0; v < vertices.size(); ++v) {
            const bool corner = vertices[v][2] < 4;
            if (o[v] > 0 && !corner) continue;
            if (o[v] > 0) vertices[v][3] = mid; // corner inside the new material's region
            vmap[v] = k++;
        }
    }
    (void)nV0;
    // merged simplex-edge pieces: positive pieces on simplex edge (i,j) form one segment
    std::vector<Edge> merged_edges;
    std::map<size_t, size_t> merged_of_piece; // positive simplex-edge piece -> merged edge slot
    if (any_pos) {
        for (size_t i = 0; i < 4; ++i)
            for (size_t j = i + 1; j < 4; ++j) {
                std::vector<size_t> pieces;
                for (size_t e = 0; e < edges.size(); ++e) {
                    if (e < nE && ecut[e].split) continue;
                    const Edge& E = edges[e];
                    if (on_simplex_edge(E) && E.m[0] == i && E.m[1] == j && edge_is_positive(e)) pieces.push_back(e);
                }
                if (pieces.empty()) continue;
                // pieces keep the direction of the simplex edge: chain from the one nobody ends at
                size_t start = NONE, end = NONE;
                for (size_t e : pieces) {
                    bool has_pred = false, has_succ = false;
                    for (size_t g : pieces) {
                        if (edges[g].v1 == edges[e].v0) has_pred = true;
                        if (edges[g].v0 == edges[e].v1) has_succ = true;
                    }
                    if (!has_pred) start = edges[e].v0;
                    if (!has_succ) end = edges[e].v1;
                }
                for (size_t e : pieces) merged_of_piece[e] = merged_edges.size();
                merged_edges.push_back({start, end, {i, j, mid}});
            }
    }
    // edge survival: positive edges vanish (simplex-edge pieces are replaced by merged edges)
    std::vector<size_t> emap(edges.size(), NONE);
    std::vector<Edge> new_edges;
    for (size_t e = 0; e < edges.size(); ++e) {
        if (e < nE && ecut[e].split) continue;
        if (any_pos && edge_is_positive(e)) continue;
        emap[e] = new_edges.size();
        new_edges.push_back(edges[e]);
    }
    const size_t merged_base = new_edges.size();
    for (auto& E : merged_edges) new_edges.push_back(E);
    auto map_edge = [&](size_t e) {
        auto it = merged_of_piece.find(e);
        return it != merged_of_piece.end() ? merged_base + it->second : emap[e];
    };
    for (auto& E : new_edges) {
        E.v0 = vmap[E.v0];
        E.v1 = vmap[E.v1];
    }
    // cells: negative / duplicate cells survive, negative sub-cells follow, the new cell is last
    std::vector<size_t> cmap(nC, NONE);
    std::vector<Cell> new_cells;
    for (size_t c = 0; c < nC; ++c)
        if (cstat[c] == C_NEG || cstat[c] == C_ZERO) {
            cmap[c] = new_cells.size();
            new_cells.push_back(Cell{{}, cells[c].material});
        }
    std::vector<size_t> neg_cell_id(nC, NONE);
    for (size_t c = 0; c < nC; ++c)
        if (cstat[c] == C_SPLIT) {
            neg_cell_id[c] = new_cells.size();
            new_cells.push_back(Cell{{}, cells[c].material});
        }
    size_t new_cell = NONE;
    if (any_pos) {
        new_cell = new_cells.size();
        new_cells.push_back(Cell{{}, mid});
    }
    // which cell lies on a given side of a surviving face
    auto side_cell = [&](size_t f, size_t old_cell, bool part_is_positive) -> size_t {
        // old_cell: the cell recorded on that side before this insertion
        if (old_cell == NONE) return NONE;
        switch (cstat[old_cell]) {
        case C_NEG:
        case C_ZERO: return cmap[old_cell];
        case C_POS: return new_cell;
        default: return part_is_positive ? new_cell : neg_cell_id[old_cell];
        }
    };
    // faces
    std::vector<Face> new_faces;
    struct Piece
    {
        size_t f;
    };
    std::vector<std::vector<size_t>> bpieces(4); // positive boundary pieces per simplex face
    for (size_t f = 0; f < n_faces_after_cut; ++f) {
        if (f < nF && fcut[f].split) continue;
        const bool is_cut_face = (f >= nF) && faces[f].bface == NONE && faces[f].neg_cell == NONE &&
                                 faces[f].pos_cell == NONE;
        Face F = faces[f];
        bool fpos = face_is_positive(f);
        if (is_cut_face) {
            // G of a split cell: find the cell
            size_t c = NONE;
            for (size_t k = 0; k < nC; ++k)
                if (cut_face_of[k] == f) c = k;
            F.pos_cell = new_cell;
            F.neg_cell = neg_cell_id[c];
        } else {
            if (any_pos && fpos) {
                if (F.bface != NONE)
                    bpieces[F.bface].push_back(f);
                continue; // interior of the new cell, or replaced by the merged boundary face
            }
            const bool part_pos = false; // surviving non-positive faces border negative parts
            size_t pc = side_cell(f, F.pos_cell, part_pos), nc = side_cell(f, F.neg_cell, part_pos);
            // a face lying in M = current (all zero) may border a positive cell
            F.pos_cell = pc;
            F.neg_cell = nc;
            if (F.bface == NONE && F.pos_cell != NONE && F.neg_cell != NONE &&
                new_cells[F.pos_cell].material < new_cells[F.neg_cell].material) {
                // keep "later material on the positive side": flip the face
                std::swap(F.pos_cell, F.neg_cell);
                std::reverse(F.verts.begin(), F.verts.end());
                // edges[k] joins verts[k], verts[k+1]: reversed loop uses the edges in reverse, shifted
                std::vector<size_t> re(F.edges.size());
                const size_t n = F.edges.size();
                for (size_t k = 0; k < n; ++k) re[k] = F.edges[(2 * n - 2 - k) % n];
                F.edges = re;
            }
        }
        for (size_t& v : F.verts) v = vmap[v];
        for (size_t& e : F.edges) e = map_edge(e);
        new_faces.push_back(F);
    }
    // merged boundary faces (i, M)
    for (size_t i = 0; i < 4; ++i) {
        if (bpieces[i].empty()) continue;
        // directed boundary edges of the union: edges of the pieces that survive (possibly merged)
        std::vector<std::array<size_t, 3>> segs; // (from, to, edge) in new numbering
        for (size_t f : bpieces[i]) {
            const Face& F = faces[f];
            const size_t n = F.verts.size();
            for (size_t k = 0; k < n; ++k) {
                size_t e = F.edges[k];
                size_t ne = map_edge(e);
                if (ne == NONE) continue; // interior edge between two pieces
                const Edge& E = new_edges[ne];
                // direction of the (possibly merged) edge along this piece's loop
                const Edge& OE = edges[e];
                bool forward = (OE.v0 == F.verts[k]);
                size_t from = forward ? E.v0 : E.v1, to = forward ? E.v1 : E.v0;
                bool dup = false;
                for (auto& s : segs)
                    if (s[2] == ne) dup = true; // merged edge contributed by several pieces
                if (!dup) segs.push_back({from, to, ne});
            }
        }
        Face Fm;
        Fm.bface = i;
        Fm.pos_cell = NONE;
        Fm.neg_cell = new_cell;
        size_t cur = segs[0][0];
        std::vector<bool> used(segs.size(), false);
        for (size_t step = 0; step < segs.size(); ++step) {
            size_t pick = NONE;
            for (size_t k = 0; k < segs.size(); ++k)
                if (!used[k] && segs[k][0] == cur) {
                    pick = k;
                    break;
                }
            if (pick == NONE) throw std::runtime_error("simplicial_arrangement(oracle): open merged boundary loop");
            used[pick] = true;
            Fm.verts.push_back(cur);
            Fm.edges.push_back(segs[pick][2]);
            cur = segs[pick][1];
        }
        if (cur != segs[0][0])
            throw std::runtime_error("simplicial_arrangement(oracle): merged boundary loop not closed");
        new_faces.push_back(Fm);
    }
    // cell face lists: faces in final order
    for (size_t f = 0; f < new_faces.size(); ++f) {
        if (new_faces[f].pos_cell != NONE) new_cells[new_faces[f].pos_cell].faces.push_back(f);
        new_cells[new_faces[f].neg_cell].faces.push_back(f);
    }
    // vertices
    std::vector<std::array<size_t, 4>> new_vertices;
    for (size_t v = 0; v < vertices.size(); ++v)
        if (vmap[v] != NONE) new_vertices.push_back(vertices[v]);
    vertices.swap(new_vertices);
    edges.swap(new_edges);
    faces.swap(new_faces);
    cells.swap(new_cells);
    return duplicate_of;
}

} // namespace

simplicial_arrangement::MaterialInterface<3> compute_material_interface_general(
    const std::vector<std::array<double, 4>>& input)
{
    using simplicial_arrangement::MaterialInterface;
    if (input.empty()) throw std::runtime_error("simplicial_arrangement(oracle): no material");
    std::vector<std::array<double, 4>> mats(4, std::array<double, 4>{0, 0, 0, 0});
    mats.insert(mats.end(), input.begin(), input.end());
    MIComplex cx;
    cx.init();
    MaterialInterface<3> out;
    out.unique_material_indices = {0, 1, 2, 3, 4};
    out.unique_materials = {{0}, {1}, {2}, {3}, {4}};
    bool has_dup = false;
    for (size_t mid = 5; mid < mats.size(); ++mid) {
        size_t dup = cx.add_material(mats, mid);
        if (dup == NONE) {
            out.unique_material_indices.push_back(out.unique_materials.size());
            out.unique_materials.push_back({mid});
        } else {
            has_dup = true;
            size_t g = out.unique_material_indices[dup];
            out.unique_material_indices.push_back(g);
            out.unique_materials[g].push_back(mid);
        }
    }
    if (!has_dup) {
        out.unique_material_indices.clear();
        out.unique_materials.clear();
    }
    out.vertices = cx.vertices;
    out.faces.resize(cx.faces.size());
    for (size_t f = 0; f < cx.faces.size(); ++f) {
        const auto& F = cx.faces[f];
        out.faces[f].vertices = F.verts;
        out.faces[f].positive_material_label = (F.pos_cell == NONE) ? F.bface : cx.cells[F.pos_cell].material;
        out.faces[f].negative_material_label = cx.cells[F.neg_cell].material;
    }
    out.cells.resize(cx.cells.size());
    for (size_t c = 0; c < cx.cells.size(); ++c) {
        out.cells[c].faces = cx.cells[c].faces;
        out.cells[c].material_label = cx.cells[c].material;
    }
    return out;
}

// ---- lookup: two materials (sign pattern of m0 - m1 at the corners); three materials go through
// the general algorithm in this restatement (identical results)
static int mi_key_2(const double* a, const double* b)
{
    int key = 0;
    for (int i = 0; i < 4; ++i) {
        if (a[i] == b[i]) return -1;
        if (a[i] > b[i]) key |= 1 << i;
    }
    return key;
}

void generate_mi_tables(std::map<int, simplicial_arrangement::MaterialInterface<3>>& mi2,
    std::map<int, simplicial_arrangement::MaterialInterface<3>>& mi3)
{
    (void)mi3;
    for (int key = 0; key < 16; ++key) {
        std::array<double, 4> a, b;
        for (int i = 0; i < 4; ++i) {
            a[i] = 0.25 * i;
            b[i] = a[i] + (((key >> i) & 1) ? -1.0 : 1.0) * (0.5 + 0.125 * i);
        }
        mi2[key] = compute_material_interface_general({a, b});
    }
}

int mi_lookup_key_2(const double* a, const double* b)
{
    return mi_key_2(a, b);
}

} // namespace sa_oracle
