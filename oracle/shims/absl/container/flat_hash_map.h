// ORACLE shim (test infrastructure, NOT product code).
// abseil (/root/reference/cmake/absl.cmake:8-9) is not vendored; the reference only needs
// try_emplace/find/erase/reserve/at/operator[] with integer and std::array keys, which
// std::unordered_map provides given a hash for std::array.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <unordered_map>
#include <unordered_set>
#include <utility>
namespace absl {
namespace shim_detail {
template <typename T>
struct Hash
{
    size_t operator()(const T& v) const { return std::hash<T>()(v); }
};
template <typename T, size_t N>
struct Hash<std::array<T, N>>
{
    size_t operator()(const std::array<T, N>& a) const
    {
        uint64_t h = 0x9e3779b97f4a7c15ULL;
        for (const T& x : a) {
            h ^= uint64_t(x) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
            h *= 0xff51afd7ed558ccdULL;
            h ^= h >> 32;
        }
        return size_t(h);
    }
};
template <typename A, typename B>
struct Hash<std::pair<A, B>>
{
    size_t operator()(const std::pair<A, B>& p) const
    {
        return Hash<std::array<uint64_t, 2>>()({uint64_t(p.first), uint64_t(p.second)});
    }
};
} // namespace shim_detail
template <typename K, typename V>
using flat_hash_map = std::unordered_map<K, V, shim_detail::Hash<K>>;
template <typename K>
using flat_hash_set = std::unordered_set<K, shim_detail::Hash<K>>;
} // namespace absl
