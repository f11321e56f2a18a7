// ORACLE shim: src/io.h:9 includes nlohmann/json.hpp but uses nothing from it; io.cpp is not built.
#pragma once
