// ORACLE (test infrastructure, NOT product code).
// Restatement of the lookup-table switches of qnzhou/simplicial_arrangement used at
// /root/reference/app/implicit_arrangement.cpp:29-42, app/material_interface.cpp:32-45,
// src/implicit_arrangement.cpp:276-280 and src/material_interface.cpp:320-324.
// Upstream loads a precomputed msgpack blob; here the tables are generated on first load by
// running the restated general algorithm on witness inputs of every sign configuration.
#pragma once
namespace simplicial_arrangement {
enum LookupTableType { ARRANGEMENT = 1, MATERIAL_INTERFACE = 2, BOTH = 3 };
bool load_lookup_table(LookupTableType type = BOTH);
void enable_lookup_table();
void disable_lookup_table();
bool lookup_table_enabled();
} // namespace simplicial_arrangement
