// ORACLE shim (test infrastructure): the un-vendored qnzhou/MshIO@main (cmake/mshio.cmake:6-7) as far as the
// reference's src/msh_io.h dereferences it.  The data structures and the MSH 4.1 serialiser are the product's
// (host/msh41.h): compiling the reference's io.cpp + msh_io.h against this shim checks that the product's
// save_result_msh fills the same blocks, tags and attributes as the reference does; the byte layout itself is
// MshIO's published format restated (parity with the real library unpinned).
#pragma once
#include "../../../robust-implicit-surface-networks_b200/host/msh41.h"

#include <fstream>
#include <stdexcept>

namespace mshio {
using MshSpec = rin_host::msh41::MshSpec;
using NodeBlock = rin_host::msh41::NodeBlock;
using ElementBlock = rin_host::msh41::ElementBlock;
using Data = rin_host::msh41::Data;
inline void validate_spec(const MshSpec&) {}
inline void save_msh(std::ostream& out, const MshSpec& spec)
{
    rin_host::msh41::write(out, spec);
}
inline void save_msh(const std::string& filename, const MshSpec& spec)
{
    std::ofstream out(filename.c_str(), std::ios::binary);
    save_msh(out, spec);
}
inline MshSpec load_msh(const std::string&)
{
    throw std::runtime_error("mshio shim: load_msh is not available");
}
inline MshSpec load_msh(std::istream&)
{
    throw std::runtime_error("mshio shim: load_msh is not available");
}
} // namespace mshio
