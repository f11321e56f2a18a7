// Minimal SIMT emulation for host-side testing of warp-cooperative device code: 32 host threads are
// the lanes of one warp; ballots / shuffles / __syncwarp are barrier-synchronised exchanges.
// Test infrastructure only (tests/test_simt_emulation.py); TSAN-friendly (pthread barriers).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <pthread.h>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __noinline__ inline // after every system header: libstdc++ spells the attribute the same way
#ifndef __restrict__
#define __restrict__
#endif

inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
#include <algorithm>
using std::max;
using std::min;
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }

namespace simt {
// Only __syncwarp() orders memory between lanes (a pthread barrier, visible to ThreadSanitizer).
// Ballots and shuffles exchange their operands through RELAXED atomics: like the hardware intrinsics they
// converge the warp but do not order other memory, so a missing __syncwarp() shows up as a TSAN race.
struct Warp
{
    pthread_barrier_t bar;
    uint64_t val[2][32];
    uint64_t tag[2][32];
};
extern Warp g_warp;
extern thread_local int t_lane;
extern thread_local uint64_t t_seq;
inline void wait() { pthread_barrier_wait(&g_warp.bar); }
// every lane posts v; returns after all lanes posted (double-buffered by sequence parity)
inline void post_and_converge(uint64_t v, uint64_t out[32])
{
    const uint64_t seq = ++t_seq;
    const int p = (int)(seq & 1);
    __atomic_store_n(&g_warp.val[p][t_lane], v, __ATOMIC_RELAXED);
    __atomic_store_n(&g_warp.tag[p][t_lane], seq, __ATOMIC_RELAXED);
    for (int i = 0; i < 32; ++i) {
        while (__atomic_load_n(&g_warp.tag[p][i], __ATOMIC_RELAXED) != seq) sched_yield();
        out[i] = __atomic_load_n(&g_warp.val[p][i], __ATOMIC_RELAXED);
    }
}
} // namespace simt

inline void __syncwarp(unsigned = 0xffffffffu) { simt::wait(); }
inline unsigned __ballot_sync(unsigned, bool p)
{
    uint64_t all[32];
    simt::post_and_converge(p ? 1 : 0, all);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= all[i] ? (1u << i) : 0u;
    return m;
}
template <class T>
inline T simt_exchange(T v, int src)
{
    uint64_t raw = 0, all[32];
    std::memcpy(&raw, &v, sizeof(T));
    simt::post_and_converge(raw, all);
    T out = v;
    if (src >= 0 && src < 32) std::memcpy(&out, &all[src], sizeof(T));
    return out;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return simt_exchange(v, src & 31); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int d) { return simt_exchange(v, simt::t_lane - d); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int d) { return simt_exchange(v, simt::t_lane ^ d); }
