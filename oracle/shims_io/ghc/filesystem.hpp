// ORACLE shim (test infrastructure): the reference's io.cpp uses ghc::filesystem, an alias of std::filesystem.
#pragma once
#include <filesystem>
namespace ghc {
namespace filesystem = std::filesystem;
}
