// ORACLE (test infrastructure, NOT product code).  Placeholder: material interface restatement.
#include <simplicial_arrangement/lookup_table.h>
#include <simplicial_arrangement/simplicial_arrangement.h>
#include "ar_complex.h"
#include <map>
namespace sa_oracle {
simplicial_arrangement::MaterialInterface<3> compute_material_interface_general(
    const std::vector<std::array<double, 4>>&)
{
    throw std::runtime_error("material interface oracle: not implemented yet");
}
void generate_mi_tables(std::map<int, simplicial_arrangement::MaterialInterface<3>>&,
    std::map<int, simplicial_arrangement::MaterialInterface<3>>&)
{}
} // namespace sa_oracle
namespace simplicial_arrangement {
MaterialInterface<3> compute_material_interface(const std::vector<Material<double, 3>>& m)
{
    return sa_oracle::compute_material_interface_general(m);
}
} // namespace simplicial_arrangement
