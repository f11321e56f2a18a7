// ORACLE (test infrastructure, NOT product code): named result vectors handed to Python via ctypes.
#pragma once
#include <simplicial_arrangement/simplicial_arrangement.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

struct ResultBag
{
    std::map<std::string, std::vector<int64_t>> i64;
    std::map<std::string, std::vector<double>> f64;
    std::string error;
    std::vector<simplicial_arrangement::Arrangement<3>> cuts_ia;
    std::vector<simplicial_arrangement::MaterialInterface<3>> cuts_mi;
    std::vector<int64_t> cut_index;
};
