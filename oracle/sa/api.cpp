// ORACLE (test infrastructure, NOT product code).
//
// Entry points + lookup tables of the restated simplicial_arrangement library.
// Upstream semantics restated (un-vendored source; call sites in /root/reference):
//  * load_lookup_table()/enable/disable are process-global switches
//    (app/implicit_arrangement.cpp:29-42, src/implicit_arrangement.cpp:276-280);
//  * IA tables cover 1 and 2 planes, indexed by the vertex signs (outer index) and, for two
//    planes, by the order of the two crossing points on every simplex edge crossed by both
//    (inner index); any zero sign or coincident crossing falls back to the general algorithm.
// Tables are generated here from the general algorithm on witness inputs; the generator
// asserts that several witnesses of one key give identical combinatorics.
#include <simplicial_arrangement/lookup_table.h>
#include <simplicial_arrangement/simplicial_arrangement.h>

#include "ar_complex.h"
#include "exact_arith.h"

#include <cstdint>
#include <map>
#include <string>

namespace sa_oracle {

static inline int sgn(double x)
{
    return x > 0 ? 1 : (x < 0 ? -1 : 0);
}

int ia_key_1(const double* p)
{
    int key = 0;
    for (int i = 0; i < 4; ++i) {
        if (p[i] == 0) return -1;
        if (p[i] > 0) key |= 1 << i;
    }
    return key;
}

int ia_key_2(const double* p0, const double* p1)
{
    int k0 = ia_key_1(p0), k1 = ia_key_1(p1);
    if (k0 < 0 || k1 < 0) return -1;
    int key = k0 | (k1 << 4);
    int e = 0;
    for (int a = 0; a < 4; ++a)
        for (int b = a + 1; b < 4; ++b, ++e) {
            bool c0 = ((k0 >> a) & 1) != ((k0 >> b) & 1);
            bool c1 = ((k1 >> a) & 1) != ((k1 >> b) & 1);
            if (!(c0 && c1)) continue;
            double m[4] = {p0[a], p0[b], p1[a], p1[b]};
            int s = det_sign(2, m, &stats().exact_fallbacks);
            if (s == 0) return -1;
            if (s > 0) key |= 1 << (8 + e);
        }
    return key;
}

static bool same_combinatorics(const simplicial_arrangement::Arrangement<3>& a,
    const simplicial_arrangement::Arrangement<3>& b)
{
    if (a.vertices != b.vertices || a.faces.size() != b.faces.size() ||
        a.cells.size() != b.cells.size())
        return false;
    for (size_t i = 0; i < a.faces.size(); ++i) {
        if (a.faces[i].vertices != b.faces[i].vertices ||
            a.faces[i].supporting_plane != b.faces[i].supporting_plane ||
            a.faces[i].positive_cell != b.faces[i].positive_cell ||
            a.faces[i].negative_cell != b.faces[i].negative_cell)
            return false;
    }
    for (size_t i = 0; i < a.cells.size(); ++i)
        if (a.cells[i].faces != b.cells[i].faces) return false;
    return true;
}

struct Tables
{
    bool ia_loaded = false, mi_loaded = false, enabled = true;
    std::map<int, simplicial_arrangement::Arrangement<3>> ia1, ia2;
    std::map<int, simplicial_arrangement::MaterialInterface<3>> mi2, mi3;
};
static Tables g_tables;

static uint64_t splitmix(uint64_t& s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static double urand(uint64_t& s)
{
    return (splitmix(s) >> 11) * (1.0 / 9007199254740992.0);
}

static void generate_ia_tables()
{
    uint64_t seed = 20220818;
    for (int key = 0; key < 16; ++key) {
        std::array<double, 4> p;
        for (int i = 0; i < 4; ++i) p[i] = ((key >> i) & 1 ? 1.0 : -1.0) * (0.5 + urand(seed));
        g_tables.ia1[key] = compute_arrangement_general({p});
    }
    std::map<int, int> hits;
    for (int it = 0; it < 400000; ++it) {
        std::array<double, 4> p0, p1;
        int outer = int(splitmix(seed) & 255);
        for (int i = 0; i < 4; ++i) {
            p0[i] = ((outer >> i) & 1 ? 1.0 : -1.0) * (0.02 + urand(seed));
            p1[i] = ((outer >> (4 + i)) & 1 ? 1.0 : -1.0) * (0.02 + urand(seed));
        }
        int key = ia_key_2(p0.data(), p1.data());
        if (key < 0) continue;
        int& h = hits[key];
        if (h == 0) {
            g_tables.ia2[key] = compute_arrangement_general({p0, p1});
        } else if (h < 8) {
            if (!same_combinatorics(g_tables.ia2[key], compute_arrangement_general({p0, p1})))
                throw std::runtime_error("simplicial_arrangement(oracle): IA 2-plane key " +
                                         std::to_string(key) + " is not a complete invariant");
        }
        ++h;
    }
}

void generate_mi_tables(std::map<int, simplicial_arrangement::MaterialInterface<3>>& mi2,
    std::map<int, simplicial_arrangement::MaterialInterface<3>>& mi3);
int mi_lookup_key_2(const double* a, const double* b);

} // namespace sa_oracle

namespace simplicial_arrangement {

EngineStats& engine_stats()
{
    return sa_oracle::stats();
}

bool load_lookup_table(LookupTableType type)
{
    auto& t = sa_oracle::g_tables;
    if ((type & ARRANGEMENT) && !t.ia_loaded) {
        sa_oracle::generate_ia_tables();
        t.ia_loaded = true;
    }
    if ((type & MATERIAL_INTERFACE) && !t.mi_loaded) {
        sa_oracle::generate_mi_tables(t.mi2, t.mi3);
        t.mi_loaded = true;
    }
    return true;
}
void enable_lookup_table()
{
    sa_oracle::g_tables.enabled = true;
}
void disable_lookup_table()
{
    sa_oracle::g_tables.enabled = false;
}
bool lookup_table_enabled()
{
    return sa_oracle::g_tables.enabled;
}

Arrangement<3> compute_arrangement(const std::vector<Plane<double, 3>>& planes)
{
    auto& t = sa_oracle::g_tables;
    if (t.enabled && t.ia_loaded) {
        if (planes.size() == 1) {
            int key = sa_oracle::ia_key_1(planes[0].data());
            if (key >= 0) {
                ++sa_oracle::stats().lookups;
                return t.ia1[key];
            }
        } else if (planes.size() == 2) {
            int key = sa_oracle::ia_key_2(planes[0].data(), planes[1].data());
            if (key >= 0) {
                auto it = t.ia2.find(key);
                if (it != t.ia2.end()) {
                    ++sa_oracle::stats().lookups;
                    return it->second;
                }
            }
        }
    }
    ++sa_oracle::stats().general;
    return sa_oracle::compute_arrangement_general(planes);
}

MaterialInterface<3> compute_material_interface(const std::vector<Material<double, 3>>& materials)
{
    auto& t = sa_oracle::g_tables;
    // upstream: 2 materials -> first table, 3 materials -> secondary table; here only the first
    // table is materialised, the 3-material case runs the general algorithm (same results)
    if (t.enabled && t.mi_loaded && materials.size() == 2) {
        int key = sa_oracle::mi_lookup_key_2(materials[0].data(), materials[1].data());
        if (key >= 0) {
            ++sa_oracle::stats().lookups;
            return t.mi2[key];
        }
    }
    ++sa_oracle::stats().general;
    return sa_oracle::compute_material_interface_general(materials);
}

} // namespace simplicial_arrangement
