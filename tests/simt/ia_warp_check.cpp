// Host check of the warp-cooperative plane insertion (csrc/ia_complex_warp.cuh) and iso record writer
// (csrc/iso_record.cuh) against the serial versions the GPU parity tests pin to the oracle.
// 32 threads play the lanes of one warp.  Usage: ia_warp_check [cases] [seed]
#include "simt_emul.h"
#include "../../robust-implicit-surface-networks_b200/csrc/iso_record.cuh"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

namespace simt {
Warp g_warp;
thread_local int t_lane = 0;
thread_local uint64_t t_seq = 0;
}
using namespace rin;

struct Case
{
    int k;
    double v[8][4];
};

template <class Caps>
static int compare(const IAComplex<Caps>& a, const IAComplex<Caps>& b, const char* what, int ci)
{
#define CMP(field)                                                                        \
    if (a.field != b.field) {                                                             \
        std::printf("case %d (%s): %s differs: %d vs %d\n", ci, what, #field, (int)a.field, (int)b.field); \
        return 1;                                                                         \
    }
    CMP(err);
    if (a.err) return 0;
    CMP(np); CMP(nv); CMP(ne); CMP(nf); CMP(nc); CMP(nfe); CMP(ncf); CMP(n_groups); CMP(has_coplanar); CMP(n_exact);
    for (int v = 0; v < a.nv; ++v)
        for (int k = 0; k < 3; ++k) CMP(vp[v][k]);
    for (int e = 0; e < a.ne; ++e) { CMP(ev0[e]); CMP(ev1[e]); CMP(ep0[e]); CMP(ep1[e]); }
    for (int f = 0; f < a.nf; ++f) {
        CMP(foff[f]); CMP(flen[f]); CMP(fplane[f]); CMP(fpos[f]); CMP(fneg[f]);
        for (int k = 0; k < a.flen[f]; ++k) { CMP(fv[a.foff[f] + k]); CMP(fe[a.foff[f] + k]); }
    }
    for (int c = 0; c < a.nc; ++c) {
        CMP(coff[c]); CMP(clen[c]);
        for (int k = 0; k < a.clen[c]; ++k) CMP(cf[a.coff[c] + k]);
    }
    for (int p = 0; p < a.np; ++p) CMP(upi[p]);
#undef CMP
    return 0;
}

template <class Caps>
struct Shared
{
    std::vector<Case> cases;
    IAComplex<Caps> cx;
    IAWarpScratch<Caps> sc;
    std::vector<IAComplex<Caps>> results;
    std::vector<std::vector<uint32_t>> records;
};

template <class Caps>
static void* lane_main(void* arg)
{
    auto* pr = static_cast<std::pair<Shared<Caps>*, int>*>(arg);
    Shared<Caps>& S = *pr->first;
    const int lane = pr->second;
    simt::t_lane = lane;
    for (size_t ci = 0; ci < S.cases.size(); ++ci) {
        const Case& C = S.cases[ci];
        if (lane == 0) S.cx.init();
        __syncwarp();
        for (int j = 0; j < C.k; ++j) warp_insert(S.cx, S.sc, C.v[j], lane);
        __syncwarp();
        if (!S.cx.err) {
            WarpIso<Caps> iso;
            iso.run(S.cx, lane);
            if (lane == 0) S.records[ci].assign(iso.size_bytes() / 4, 0xdeadbeefu);
            __syncwarp();
            iso.write(S.cx, S.records[ci].data(), lane);
        }
        __syncwarp();
        if (lane == 0) S.results[ci] = S.cx;
        __syncwarp();
    }
    return nullptr;
}

template <class Caps>
static int run(int n_cases, unsigned seed, int maxk, const char* name)
{
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    Shared<Caps>* S = new Shared<Caps>();
    for (int ci = 0; ci < n_cases; ++ci) {
        Case C;
        C.k = 1 + (int)(rng() % maxk);
        const int flavour = (int)(rng() % 8);
        for (int j = 0; j < C.k; ++j) {
            for (int c = 0; c < 4; ++c) {
                double x = U(rng);
                if (flavour == 1) x = (double)((int)(rng() % 5) - 2);          // small integers: ties, zeros
                if (flavour == 2 && (rng() % 3) == 0) x = 0.0;                // planes through corners
                if (flavour == 3) x = std::round(x * 4.0) / 4.0;              // coarse lattice
                C.v[j][c] = x;
            }
            if (flavour == 4 && j > 0 && (rng() % 2) == 0) {                   // coincident planes (scaled copies)
                const double s = (rng() % 2) ? 2.0 : -0.5;
                for (int c = 0; c < 4; ++c) C.v[j][c] = s * C.v[j - 1][c];
            }
            if (flavour == 5 && j > 0)                                        // nearly coincident
                for (int c = 0; c < 4; ++c) C.v[j][c] = C.v[0][c] + (c == j % 4 ? 1e-15 : 0.0);
        }
        S->cases.push_back(C);
    }
    S->results.resize(n_cases);
    S->records.resize(n_cases);
    std::memset(&simt::g_warp, 0, sizeof(simt::g_warp)); // lanes restart their sequence numbers at 0
    pthread_barrier_init(&simt::g_warp.bar, nullptr, 32);
    pthread_t th[32];
    std::pair<Shared<Caps>*, int> args[32];
    for (int l = 0; l < 32; ++l) {
        args[l] = {S, l};
        pthread_create(&th[l], nullptr, lane_main<Caps>, &args[l]);
    }
    for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&simt::g_warp.bar);

    int bad = 0, n_err[3] = {0, 0, 0}, cap_diff = 0;
    long exact = 0;
    for (int ci = 0; ci < n_cases && bad < 5; ++ci) {
        const Case& C = S->cases[ci];
        IAComplex<Caps>* ser = new IAComplex<Caps>();
        ser->init();
        for (int j = 0; j < C.k; ++j) ser->insert(C.v[j]);
        const IAComplex<Caps>& wr = S->results[ci];
        n_err[ser->err]++;
        exact += ser->n_exact;
        if (ser->err == 1 && wr.err == 0) { // the serial capacity check is conservative
            ++cap_diff;
            delete ser;
            continue;
        }
        if (compare(*ser, wr, name, ci)) {
            ++bad;
            delete ser;
            continue;
        }
        if (!ser->err) {
            IsoScan<Caps> iso;
            iso.run(*ser);
            std::vector<uint32_t> rec(iso.size_bytes() / 4, 0xdeadbeefu);
            iso.write(*ser, rec.data());
            if (rec != S->records[ci]) {
                std::printf("case %d (%s): iso record differs (%zu vs %zu words)\n", ci, name, rec.size(),
                    S->records[ci].size());
                ++bad;
            }
        }
        delete ser;
    }
    std::printf("%s: %d cases, ok=%d degenerate=%d capacity=%d (serial-only capacity %d), exact fallbacks %ld, mismatches %d\n",
        name, n_cases, n_err[0], n_err[2], n_err[1], cap_diff, exact, bad);
    delete S;
    return bad;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? std::atoi(argv[1]) : 2000;
    const unsigned seed = argc > 2 ? (unsigned)std::atoi(argv[2]) : 1u;
    int bad = 0;
    bad += run<IACapsSmall>(n, seed, 6, "small tier caps (k<=6: more than the tier holds -> capacity path)");
    bad += run<IACapsMid>(n / 4 + 1, seed + 3, 10, "mid tier caps (k<=10)");
    bad += run<IACaps>(n / 4 + 1, seed + 7, 8, "big tier caps (k<=8)");
    return bad ? 1 : 0;
}
