// MSH 4.1 (Gmsh) binary writer for the three result files of save_result_msh
// (/root/reference/src/io.cpp:374-551 through src/msh_io.h:199-330, which delegates to the un-vendored
// qnzhou/MshIO@main, cmake/mshio.cmake:6-7).  The structures mirror what msh_io.h fills; the byte layout follows
// the published MSH 4.1 format (section headers in ASCII, payload in binary, data-size 8):
//   $MeshFormat  "4.1 1 8", the int 1 in binary (endianness probe)
//   $Nodes       4 size_t (blocks, nodes, min tag, max tag); per block 3 int (dim, tag, parametric) + size_t count,
//                the node tags (size_t), then x y z (double) per node
//   $Elements    4 size_t; per block 3 int (dim, tag, element type) + size_t count, then per element
//                tag + node tags (size_t)
//   $NodeData / $ElementData  ASCII tag header (string / real / integer tags), then per entry int tag + doubles
// Byte parity with MshIO itself is unpinned (the library is absent here); tests read the files back per this layout.
#pragma once
#include <cstdint>
#include <ostream>
#include <string>
#include <vector>

namespace rin_host {
namespace msh41 {

struct NodeBlock
{
    int entity_dim = 0, entity_tag = 0, parametric = 0;
    size_t num_nodes_in_block = 0;
    std::vector<size_t> tags;
    std::vector<double> data; // x y z per node
};
struct Nodes
{
    size_t num_entity_blocks = 0, num_nodes = 0, min_node_tag = 0, max_node_tag = 0;
    std::vector<NodeBlock> entity_blocks;
};
struct ElementBlock
{
    int entity_dim = 0, entity_tag = 0, element_type = 0;
    size_t num_elements_in_block = 0;
    std::vector<size_t> data; // element tag + node tags per element
};
struct Elements
{
    size_t num_entity_blocks = 0, num_elements = 0, min_element_tag = 0, max_element_tag = 0;
    std::vector<ElementBlock> entity_blocks;
};
struct DataHeader
{
    std::vector<std::string> string_tags;
    std::vector<double> real_tags;
    std::vector<int> int_tags;
};
struct DataEntry
{
    size_t tag = 0;
    std::vector<double> data;
};
struct Data
{
    DataHeader header;
    std::vector<DataEntry> entries;
};
struct MeshFormat
{
    std::string version = "4.1";
    int file_type = 1; // binary
    int data_size = sizeof(size_t);
};
struct MshSpec
{
    MeshFormat mesh_format;
    Nodes nodes;
    Elements elements;
    std::vector<Data> node_data, element_data;
};

namespace detail {
template <typename T>
inline void put(std::ostream& out, const T& v)
{
    out.write(reinterpret_cast<const char*>(&v), sizeof(T));
}
inline void write_data(std::ostream& out, const char* section, const Data& d)
{
    out << "$" << section << "\n";
    out << d.header.string_tags.size() << "\n";
    for (const auto& s : d.header.string_tags) out << "\"" << s << "\"\n";
    out << d.header.real_tags.size() << "\n";
    for (double r : d.header.real_tags) out << r << "\n";
    out << d.header.int_tags.size() << "\n";
    for (int i : d.header.int_tags) out << i << "\n";
    for (const auto& e : d.entries) {
        put<int32_t>(out, (int32_t)e.tag);
        out.write(reinterpret_cast<const char*>(e.data.data()), (std::streamsize)(e.data.size() * sizeof(double)));
    }
    out << "\n$End" << section << "\n";
}
} // namespace detail

inline void write(std::ostream& out, const MshSpec& m)
{
    using detail::put;
    out.precision(16);
    out << "$MeshFormat\n" << m.mesh_format.version << " " << m.mesh_format.file_type << " " << m.mesh_format.data_size
        << "\n";
    put<int>(out, 1);
    out << "\n$EndMeshFormat\n";
    if (m.nodes.num_entity_blocks) {
        out << "$Nodes\n";
        put<size_t>(out, m.nodes.num_entity_blocks);
        put<size_t>(out, m.nodes.num_nodes);
        put<size_t>(out, m.nodes.min_node_tag);
        put<size_t>(out, m.nodes.max_node_tag);
        for (const auto& b : m.nodes.entity_blocks) {
            put<int>(out, b.entity_dim);
            put<int>(out, b.entity_tag);
            put<int>(out, b.parametric);
            put<size_t>(out, b.num_nodes_in_block);
            out.write(reinterpret_cast<const char*>(b.tags.data()), (std::streamsize)(b.tags.size() * sizeof(size_t)));
            out.write(reinterpret_cast<const char*>(b.data.data()), (std::streamsize)(b.data.size() * sizeof(double)));
        }
        out << "\n$EndNodes\n";
    }
    if (m.elements.num_entity_blocks) {
        out << "$Elements\n";
        put<size_t>(out, m.elements.num_entity_blocks);
        put<size_t>(out, m.elements.num_elements);
        put<size_t>(out, m.elements.min_element_tag);
        put<size_t>(out, m.elements.max_element_tag);
        for (const auto& b : m.elements.entity_blocks) {
            put<int>(out, b.entity_dim);
            put<int>(out, b.entity_tag);
            put<int>(out, b.element_type);
            put<size_t>(out, b.num_elements_in_block);
            out.write(reinterpret_cast<const char*>(b.data.data()), (std::streamsize)(b.data.size() * sizeof(size_t)));
        }
        out << "\n$EndElements\n";
    }
    for (const auto& d : m.node_data) detail::write_data(out, "NodeData", d);
    for (const auto& d : m.element_data) detail::write_data(out, "ElementData", d);
}

} // namespace msh41
} // namespace rin_host
