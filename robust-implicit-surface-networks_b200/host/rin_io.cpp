// See rin_io.h.  A small JSON reader (config, tet mesh), a JSON emitter reproducing nlohmann::json::dump() byte
// for byte for the value kinds the reference writes, and the population of the three MSH files.
#include "rin_io.h"

#include "msh41.h"

#include <charconv>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <tuple>

namespace rin_host {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// reader
// ---------------------------------------------------------------------------------------------------------------
struct JValue
{
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    bool is_int = false;
    unsigned long long u = 0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;

    const JValue* find(const std::string& k) const
    {
        for (const auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const JValue& at(const std::string& k) const
    {
        const JValue* v = find(k);
        if (!v) throw std::runtime_error("config: key '" + k + "' not found");
        return *v;
    }
};

struct JParser
{
    const char* p;
    const char* e;
    void ws()
    {
        while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    }
    [[noreturn]] void bad(const char* what) { throw std::runtime_error(std::string("JSON parse error: ") + what); }
    JValue value()
    {
        ws();
        if (p >= e) bad("unexpected end");
        JValue v;
        if (*p == '{') {
            v.kind = JValue::Object;
            ++p;
            ws();
            if (p < e && *p == '}') {
                ++p;
                return v;
            }
            for (;;) {
                ws();
                if (p >= e || *p != '"') bad("expected a key");
                std::string k = string();
                ws();
                if (p >= e || *p != ':') bad("expected ':'");
                ++p;
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (p < e && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < e && *p == '}') {
                    ++p;
                    return v;
                }
                bad("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JValue::Array;
            ++p;
            ws();
            if (p < e && *p == ']') {
                ++p;
                return v;
            }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (p < e && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < e && *p == ']') {
                    ++p;
                    return v;
                }
                bad("expected ',' or ']'");
            }
        }
        if (*p == '"') {
            v.kind = JValue::String;
            v.str = string();
            return v;
        }
        if (e - p >= 4 && !strncmp(p, "true", 4)) {
            p += 4;
            v.kind = JValue::Bool;
            v.b = true;
            return v;
        }
        if (e - p >= 5 && !strncmp(p, "false", 5)) {
            p += 5;
            v.kind = JValue::Bool;
            return v;
        }
        if (e - p >= 4 && !strncmp(p, "null", 4)) {
            p += 4;
            return v;
        }
        // number
        const char* s = p;
        bool integral = true;
        if (p < e && *p == '-') ++p;
        while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
            if (*p == '.' || *p == 'e' || *p == 'E') integral = false;
            ++p;
        }
        if (p == s) bad("unexpected character");
        v.kind = JValue::Number;
        auto r = std::from_chars(s, p, v.num);
        if (r.ec != std::errc()) bad("bad number");
        if (integral && *s != '-') {
            auto r2 = std::from_chars(s, p, v.u);
            v.is_int = r2.ec == std::errc();
        }
        return v;
    }
    std::string string()
    {
        std::string out;
        ++p; // opening quote
        while (p < e && *p != '"') {
            if (*p == '\\') {
                ++p;
                if (p >= e) bad("bad escape");
                switch (*p) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': { // basic multilingual plane only (paths and labels are ASCII in practice)
                    if (e - p < 5) bad("bad \\u escape");
                    unsigned cp = 0;
                    for (int i = 1; i <= 4; ++i) {
                        const char c = p[i];
                        cp = cp * 16 + (c >= '0' && c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
                    }
                    p += 4;
                    if (cp < 0x80)
                        out += char(cp);
                    else if (cp < 0x800) {
                        out += char(0xc0 | (cp >> 6));
                        out += char(0x80 | (cp & 63));
                    } else {
                        out += char(0xe0 | (cp >> 12));
                        out += char(0x80 | ((cp >> 6) & 63));
                        out += char(0x80 | (cp & 63));
                    }
                    break;
                }
                default: out += *p;
                }
                ++p;
            } else
                out += *p++;
        }
        if (p >= e) bad("unterminated string");
        ++p;
        return out;
    }
};

JValue parse_file(const std::string& filename)
{
    std::ifstream fin(filename.c_str(), std::ios::binary);
    std::stringstream ss;
    ss << fin.rdbuf();
    const std::string text = ss.str();
    JParser ps{text.data(), text.data() + text.size()};
    return ps.value();
}

std::string resolve(const std::filesystem::path& base, const std::string& p)
{
    std::filesystem::path q(p);
    if (q.is_relative()) q = std::filesystem::absolute(base / q);
    return q.string();
}

// ---------------------------------------------------------------------------------------------------------------
// emitter: the notation of nlohmann::json::dump() (compact)
// ---------------------------------------------------------------------------------------------------------------
void put_u64(std::string& o, unsigned long long v)
{
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    o.append(buf, r.ptr);
}

// doubles: shortest digits that round-trip, laid out as nlohmann does (detail::to_chars / format_buffer):
// fixed notation while the decimal point lies within [-4, 15] digits, else d[.ddd]e[+-]XX; integral values keep ".0"
void put_double(std::string& o, double x)
{
    if (!(x == x) || x - x != 0) { // nan / inf
        o += "null";
        return;
    }
    if (x == 0) {
        o += std::signbit(x) ? "-0.0" : "0.0";
        return;
    }
    if (x < 0) {
        o += '-';
        x = -x;
    }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof sci, x, std::chars_format::scientific);
    // sci = d[.ddd]e[+-]XX
    char digits[24];
    int k = 0;
    const char* q = sci;
    for (; q < r.ptr && *q != 'e'; ++q)
        if (*q != '.') digits[k++] = *q;
    int e10 = 0;
    std::from_chars(q + 1 + (q[1] == '+'), r.ptr, e10);
    const int n = e10 + 1; // position of the decimal point relative to the first digit
    constexpr int min_exp = -4, max_exp = 15;
    if (k <= n && n <= max_exp) { // digits[000].0
        o.append(digits, k);
        o.append(size_t(n - k), '0');
        o += ".0";
    } else if (0 < n && n <= max_exp) { // dig.its
        o.append(digits, n);
        o += '.';
        o.append(digits + n, k - n);
    } else if (min_exp < n && n <= 0) { // 0.[000]digits
        o += "0.";
        o.append(size_t(-n), '0');
        o.append(digits, k);
    } else { // d[.igits]e+XX
        o += digits[0];
        if (k > 1) {
            o += '.';
            o.append(digits + 1, k - 1);
        }
        o += 'e';
        int ex = n - 1;
        o += ex < 0 ? '-' : '+';
        if (ex < 0) ex = -ex;
        if (ex < 10) o += '0';
        put_u64(o, (unsigned long long)ex);
    }
}

void put_string(std::string& o, const std::string& s)
{
    o += '"';
    for (unsigned char c : s) {
        switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\b': o += "\\b"; break;
        case '\f': o += "\\f"; break;
        case '\n': o += "\\n"; break;
        case '\r': o += "\\r"; break;
        case '\t': o += "\\t"; break;
        default:
            if (c < 0x20) {
                char buf[8];
                snprintf(buf, sizeof buf, "\\u%04x", c);
                o += buf;
            } else
                o += char(c);
        }
    }
    o += '"';
}

template <typename It>
void put_u64_array(std::string& o, It b, It e)
{
    o += '[';
    for (It i = b; i != e; ++i) {
        if (i != b) o += ',';
        put_u64(o, (unsigned long long)*i);
    }
    o += ']';
}

// a json that only ever receives push_back: `null` when nothing was pushed (src/io.cpp:169-217)
template <typename Seq, typename Fn>
void put_pushed(std::string& o, const Seq& seq, Fn each)
{
    if (seq.empty()) {
        o += "null";
        return;
    }
    o += '[';
    bool first = true;
    for (const auto& x : seq) {
        if (!first) o += ',';
        first = false;
        each(x);
    }
    o += ']';
}

struct MeshJson
{
    std::string points, faces, patches, edges, chains, corners;
    MeshJson(const std::vector<std::array<double, 3>>& pts, const std::vector<PolygonFace>& mesh_faces,
        const std::vector<std::vector<size_t>>& patch_list, const std::vector<Edge>& edge_list,
        const std::vector<std::vector<size_t>>& chain_list, const std::vector<std::vector<size_t>>& nme)
    {
        points.reserve(pts.size() * 60);
        put_pushed(points, pts, [&](const std::array<double, 3>& p) {
            points += '[';
            put_double(points, p[0]);
            points += ',';
            put_double(points, p[1]);
            points += ',';
            put_double(points, p[2]);
            points += ']';
        });
        put_pushed(faces, mesh_faces,
            [&](const PolygonFace& f) { put_u64_array(faces, f.vert_indices.begin(), f.vert_indices.end()); });
        put_pushed(patches, patch_list, [&](const std::vector<size_t>& x) { put_u64_array(patches, x.begin(), x.end()); });
        put_pushed(edges, edge_list, [&](const Edge& ed) {
            const size_t v[2] = {ed.v1, ed.v2};
            put_u64_array(edges, v, v + 2);
        });
        put_pushed(chains, chain_list, [&](const std::vector<size_t>& x) { put_u64_array(chains, x.begin(), x.end()); });
        // corners: vertices with more than two, or exactly one, incident non-manifold edges (src/io.cpp:199)
        std::vector<size_t> cs;
        for (size_t i = 0; i < nme.size(); ++i)
            if (nme[i].size() > 2 || nme[i].size() == 1) cs.push_back(i);
        if (cs.empty())
            corners = "null";
        else
            put_u64_array(corners, cs.begin(), cs.end());
    }
};

bool write_line(const std::string& filename, const std::string& line)
{
    std::ofstream fout(filename.c_str(), std::ios::binary);
    fout << line << std::endl;
    return bool(fout);
}

std::string nested(const std::vector<std::vector<size_t>>& lists)
{
    std::string o;
    put_pushed(o, lists, [&](const std::vector<size_t>& x) { put_u64_array(o, x.begin(), x.end()); });
    return o;
}

} // namespace

std::string json_number(double x)
{
    std::string o;
    put_double(o, x);
    return o;
}

Config parse_config_file(const std::string& filename)
{
    {
        std::ifstream probe(filename.c_str());
        if (!probe) throw std::runtime_error("Config file does not exist!");
    }
    const JValue data = parse_file(filename);
    Config config;
    const std::filesystem::path dir = std::filesystem::path(filename).parent_path();
    if (data.find("tetMeshFile")) {
        config.tet_mesh_file = resolve(dir, data.at("tetMeshFile").str);
        config.tet_mesh_resolution = 0;
    } else {
        config.tet_mesh_file = "";
        config.tet_mesh_resolution = (size_t)data.at("gridResolution").num;
        const JValue& bb = data.at("gridBbox");
        for (int k = 0; k < 3; ++k) {
            config.tet_mesh_bbox_min[k] = bb.arr.at(0).arr.at(k).num;
            config.tet_mesh_bbox_max[k] = bb.arr.at(1).arr.at(k).num;
        }
    }
    config.func_file = resolve(dir, data.at("funcFile").str);
    config.output_dir = resolve(dir, data.at("outputDir").str);
    config.use_lookup = data.at("useLookup").b;
    config.use_secondary_lookup = data.at("useSecondaryLookup").b;
    config.use_topo_ray_shooting = data.at("useTopoRayShooting").b;
    return config;
}

bool load_tet_mesh(const std::string& filename, std::vector<std::array<double, 3>>& pts,
    std::vector<std::array<size_t, 4>>& tets)
{
    {
        std::ifstream probe(filename.c_str());
        if (!probe) {
            std::cout << "tet mesh file not exist!" << std::endl;
            return false;
        }
    }
    const JValue data = parse_file(filename);
    const auto& jp = data.arr.at(0).arr;
    const auto& jt = data.arr.at(1).arr;
    pts.resize(jp.size());
    for (size_t j = 0; j < pts.size(); ++j)
        for (size_t k = 0; k < 3; ++k) pts[j][k] = jp[j].arr.at(k).num;
    tets.resize(jt.size());
    for (size_t j = 0; j < tets.size(); ++j)
        for (size_t k = 0; k < 4; ++k) tets[j][k] = (size_t)jt[j].arr.at(k).u;
    return true;
}

bool save_result(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<size_t>& patch_function_label, const std::vector<Edge>& edges,
    const std::vector<std::vector<size_t>>& chains, const std::vector<std::vector<size_t>>& non_manifold_edges_of_vert,
    const std::vector<std::vector<size_t>>& shells, const std::vector<std::vector<size_t>>& cells,
    const std::vector<std::vector<bool>>& cell_function_label)
{
    const MeshJson m(mesh_pts, mesh_faces, patches, edges, chains, non_manifold_edges_of_vert);
    std::string labels;
    put_pushed(labels, cell_function_label, [&](const std::vector<bool>& row) {
        labels += '[';
        for (size_t i = 0; i < row.size(); ++i) {
            if (i) labels += ',';
            labels += row[i] ? "true" : "false";
        }
        labels += ']';
    });
    std::string plabel = "[";
    put_u64_array(plabel, patch_function_label.begin(), patch_function_label.end());
    plabel += ']';
    // keys in lexicographic order (nlohmann::json objects are ordered maps)
    std::string o;
    o.reserve(m.points.size() + m.faces.size() + m.patches.size() + m.edges.size() + 4096);
    o += "{\"cells\":" + nested(cells) + ",\"cells_label\":" + labels + ",\"chains\":" + m.chains + ",\"corners\":" +
         m.corners + ",\"edges\":" + m.edges + ",\"faces\":" + m.faces + ",\"patches\":" + m.patches +
         ",\"patches_label\":" + plabel + ",\"points\":" + m.points + ",\"shells\":" + nested(shells) + "}";
    return write_line(filename, o);
}

bool save_result_MI(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<std::pair<size_t, size_t>>& patch_function_label, const std::vector<Edge>& edges,
    const std::vector<std::vector<size_t>>& chains, const std::vector<std::vector<size_t>>& non_manifold_edges_of_vert,
    const std::vector<std::vector<size_t>>& shells, const std::vector<std::vector<size_t>>& cells,
    const std::vector<size_t>& cell_function_label)
{
    const MeshJson m(mesh_pts, mesh_faces, patches, edges, chains, non_manifold_edges_of_vert);
    std::string plabel = "[[";
    for (size_t i = 0; i < patch_function_label.size(); ++i) {
        if (i) plabel += ',';
        const size_t v[2] = {patch_function_label[i].first, patch_function_label[i].second};
        put_u64_array(plabel, v, v + 2);
    }
    plabel += "]]";
    std::string clabel = "[";
    put_u64_array(clabel, cell_function_label.begin(), cell_function_label.end());
    clabel += ']';
    std::string o;
    o.reserve(m.points.size() + m.faces.size() + m.patches.size() + m.edges.size() + 4096);
    o += "{\"cells\":" + nested(cells) + ",\"cells_label\":" + clabel + ",\"chains\":" + m.chains + ",\"corners\":" +
         m.corners + ",\"edges\":" + m.edges + ",\"faces\":" + m.faces + ",\"patches\":" + m.patches +
         ",\"patches_label\":" + plabel + ",\"points\":" + m.points + ",\"shells\":" + nested(shells) + "}";
    return write_line(filename, o);
}

bool save_result_CSG(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<bool>& patch_sign_label, const std::vector<Edge>& edges,
    const std::vector<std::vector<size_t>>& chains, const std::vector<std::vector<size_t>>& non_manifold_edges_of_vert)
{
    const MeshJson m(mesh_pts, mesh_faces, patches, edges, chains, non_manifold_edges_of_vert);
    std::string plabel = "[[";
    for (size_t i = 0; i < patch_sign_label.size(); ++i) {
        if (i) plabel += ',';
        plabel += patch_sign_label[i] ? "true" : "false";
    }
    plabel += "]]";
    std::string o;
    o += "{\"chains\":" + m.chains + ",\"corners\":" + m.corners + ",\"edges\":" + m.edges + ",\"faces\":" + m.faces +
         ",\"patches\":" + m.patches + ",\"patches_label\":" + plabel + ",\"points\":" + m.points + "}";
    return write_line(filename, o);
}

bool save_timings(const std::string& filename, const std::vector<std::string>& timing_labels,
    const std::vector<double>& timings)
{
    std::map<std::string, double> sorted; // a repeated label keeps its last value, as operator[] assignment does
    for (size_t i = 0; i < timings.size(); ++i) sorted[timing_labels[i]] = timings[i];
    std::string o;
    if (sorted.empty())
        o = "null";
    else {
        o = "{";
        bool first = true;
        for (const auto& kv : sorted) {
            if (!first) o += ',';
            first = false;
            put_string(o, kv.first);
            o += ':';
            put_double(o, kv.second);
        }
        o += '}';
    }
    return write_line(filename, o);
}

bool save_statistics(const std::string& filename, const std::vector<std::string>& stats_labels,
    const std::vector<size_t>& stats)
{
    std::map<std::string, size_t> sorted;
    for (size_t i = 0; i < stats.size(); ++i) sorted[stats_labels[i]] = stats[i];
    std::string o;
    if (sorted.empty())
        o = "null";
    else {
        o = "{";
        bool first = true;
        for (const auto& kv : sorted) {
            if (!first) o += ',';
            first = false;
            put_string(o, kv.first);
            o += ':';
            put_u64(o, kv.second);
        }
        o += '}';
    }
    return write_line(filename, o);
}

// ---------------------------------------------------------------------------------------------------------------
// MSH: the three files of save_result_msh (src/io.cpp:374-551) - chains as line elements, patches and cells as
// fan-triangulated polygons, one node / element block per chain, patch or cell (src/msh_io.h:217-300)
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct MshBuilder
{
    msh41::MshSpec spec;

    void add_vertices(int dim, const std::vector<std::array<double, 3>>& v)
    {
        if (v.empty()) return;
        msh41::NodeBlock b;
        b.num_nodes_in_block = v.size();
        b.entity_dim = dim;
        b.entity_tag = (int)spec.nodes.num_entity_blocks + 1;
        const size_t tag0 = spec.nodes.max_node_tag;
        for (size_t i = 0; i < v.size(); ++i) {
            b.tags.push_back(tag0 + i + 1);
            b.data.insert(b.data.end(), v[i].begin(), v[i].end());
        }
        spec.nodes.num_entity_blocks += 1;
        spec.nodes.num_nodes += v.size();
        spec.nodes.min_node_tag = 1;
        spec.nodes.max_node_tag += v.size();
        spec.nodes.entity_blocks.push_back(std::move(b));
    }
    template <size_t K>
    void add_elements(int dim, const std::vector<std::array<size_t, K>>& el)
    {
        if (el.empty()) return;
        if (spec.nodes.num_nodes == 0) throw std::runtime_error("Please add a vertex block before adding elements.");
        const msh41::NodeBlock& vb = spec.nodes.entity_blocks.back();
        if (vb.entity_dim != dim)
            throw std::runtime_error("It seems the last added vertex block has different dimension "
                                     "than the elements you want to add.");
        msh41::ElementBlock b;
        b.entity_dim = dim;
        b.entity_tag = vb.entity_tag;
        b.element_type = dim == 1 ? 1 : (dim == 2 ? 2 : 4); // 2-node line, 3-node triangle, 4-node tet
        b.num_elements_in_block = el.size();
        const size_t voff = vb.tags.front() - 1, tag0 = spec.elements.max_element_tag;
        for (size_t i = 0; i < el.size(); ++i) {
            b.data.push_back(tag0 + i + 1);
            for (size_t j = 0; j < K; ++j) b.data.push_back(voff + el[i][j] + 1);
        }
        spec.elements.num_entity_blocks++;
        spec.elements.num_elements += el.size();
        spec.elements.min_element_tag = 1;
        spec.elements.max_element_tag += el.size();
        spec.elements.entity_blocks.push_back(std::move(b));
    }
    void add_face_attribute(const std::string& name, const std::vector<size_t>& values)
    {
        if (spec.elements.entity_blocks.empty()) throw std::runtime_error("Please add elements before adding element attributes!");
        msh41::Data d;
        d.header.string_tags = {name};
        d.header.real_tags = {0.0};
        size_t total = 0;
        for (const auto& eb : spec.elements.entity_blocks) {
            if (eb.entity_dim != 2) continue;
            for (size_t i = 0; i < eb.num_elements_in_block; ++i) {
                msh41::DataEntry en;
                en.tag = eb.data[i * 4];
                en.data.push_back(double(values[total + i]));
                d.entries.push_back(std::move(en));
            }
            total += eb.num_elements_in_block;
        }
        d.header.int_tags = {0, 1, int(total), 0, 2};
        spec.element_data.push_back(std::move(d));
    }
    bool save(const std::string& filename)
    {
        std::ofstream out(filename.c_str(), std::ios::binary);
        msh41::write(out, spec);
        return bool(out);
    }
};

// local numbering in order of first use, as the reference's vertex_map passes do
struct LocalVerts
{
    std::vector<size_t>& map;
    size_t n = 0;
    static constexpr size_t INVALID = ~size_t(0);
    explicit LocalVerts(std::vector<size_t>& m) : map(m) { std::fill(map.begin(), map.end(), INVALID); }
    void use(size_t v)
    {
        if (map[v] == INVALID) map[v] = n++;
    }
    std::vector<std::array<double, 3>> gather(const std::vector<std::array<double, 3>>& pts) const
    {
        std::vector<std::array<double, 3>> out(n);
        for (size_t j = 0; j < map.size(); ++j)
            if (map[j] != INVALID) out[map[j]] = pts[j];
        return out;
    }
};

void fan(const PolygonFace& f, LocalVerts& lv, std::vector<std::array<size_t, 3>>& tris, std::vector<size_t>* ids,
    size_t poly_id)
{
    const auto& v = f.vert_indices;
    lv.use(v[0]);
    lv.use(v[1]);
    for (size_t j = 2; j < v.size(); ++j) {
        tris.push_back({v[0], v[j - 1], v[j]});
        if (ids) ids->push_back(poly_id);
        lv.use(v[j]);
    }
}

} // namespace

bool save_result_msh(const std::string& filename, const std::vector<std::array<double, 3>>& mesh_pts,
    const std::vector<PolygonFace>& mesh_faces, const std::vector<std::vector<size_t>>& patches,
    const std::vector<Edge>& edges, const std::vector<std::vector<size_t>>& chains,
    const std::vector<std::vector<size_t>>& /*non_manifold_edges_of_vert*/, const std::vector<std::vector<size_t>>& shells,
    const std::vector<std::vector<size_t>>& cells)
{
    std::vector<size_t> vmap(mesh_pts.size());
    MshBuilder m_chains, m_patches, m_cells;
    for (const auto& chain : chains) {
        LocalVerts lv(vmap);
        for (size_t e : chain) {
            lv.use(edges[e].v1);
            lv.use(edges[e].v2);
        }
        std::vector<std::array<size_t, 2>> lines;
        for (size_t e : chain) lines.push_back({vmap[edges[e].v1], vmap[edges[e].v2]});
        m_chains.add_vertices(1, lv.gather(mesh_pts));
        m_chains.add_elements<2>(1, lines);
    }
    bool ok = m_chains.save(filename + "_chains.msh");

    std::vector<size_t> patch_ids, polygon_ids;
    for (size_t i = 0; i < patches.size(); ++i) {
        LocalVerts lv(vmap);
        std::vector<std::array<size_t, 3>> tris;
        for (size_t f : patches[i]) fan(mesh_faces[f], lv, tris, &polygon_ids, f);
        for (auto& t : tris)
            for (auto& x : t) x = vmap[x];
        m_patches.add_vertices(2, lv.gather(mesh_pts));
        m_patches.add_elements<3>(2, tris);
        patch_ids.insert(patch_ids.end(), tris.size(), i);
    }
    if (!patches.empty()) {
        m_patches.add_face_attribute("patch_id", patch_ids);
        m_patches.add_face_attribute("polygon_id", polygon_ids);
    }
    ok &= m_patches.save(filename + "_patches.msh");

    if (!patches.empty()) {
        std::vector<size_t> cell_ids;
        for (size_t i = 0; i < cells.size(); ++i) {
            LocalVerts lv(vmap);
            std::vector<std::array<size_t, 3>> tris;
            for (size_t s : cells[i])
                for (size_t half_patch : shells[s])
                    for (size_t f : patches[half_patch / 2]) fan(mesh_faces[f], lv, tris, nullptr, f);
            for (auto& t : tris)
                for (auto& x : t) x = vmap[x];
            m_cells.add_vertices(2, lv.gather(mesh_pts));
            m_cells.add_elements<3>(2, tris);
            cell_ids.insert(cell_ids.end(), tris.size(), i);
        }
        m_cells.add_face_attribute("cell_id", cell_ids);
        ok &= m_cells.save(filename + "_cells.msh");
    }
    return ok;
}

} // namespace rin_host
