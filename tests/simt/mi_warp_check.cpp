// Host check of the warp-cooperative material insertion (csrc/mi_complex_warp.cuh) against the serial
// version the GPU parity tests pin to the oracle.  32 threads play the lanes of one warp.
// Usage: mi_warp_check [cases] [seed]
#include "simt_emul.h"
#include "../../robust-implicit-surface-networks_b200/csrc/mi_complex_warp.cuh"

namespace simt {
Warp g_warp;
thread_local int t_lane = 0;
thread_local uint64_t t_seq = 0;
}
using namespace rin;

struct Case
{
    int k;
    double v[10][4];
};

template <class Caps>
static int compare(const MIComplex<Caps>& a, const MIComplex<Caps>& b, const char* what, int ci)
{
#define CMP(field)                                                                        \
    if (a.field != b.field) {                                                             \
        std::printf("case %d (%s): %s differs: %d vs %d\n", ci, what, #field, (int)a.field, (int)b.field); \
        return 1;                                                                         \
    }
    CMP(err);
    if (a.err) return 0;
    CMP(nm); CMP(nv); CMP(ne); CMP(nf); CMP(nc); CMP(nfe); CMP(n_groups); CMP(has_dup); CMP(n_exact); CMP(cur);
    const int B = a.cur;
    for (int v = 0; v < a.nv; ++v)
        for (int k = 0; k < 4; ++k) CMP(vm[v][k]);
    for (int e = 0; e < a.ne; ++e) { CMP(ev0[e]); CMP(ev1[e]); CMP(em[e][0]); CMP(em[e][1]); CMP(em[e][2]); }
    for (int f = 0; f < a.nf; ++f) {
        CMP(foff[B][f]); CMP(flen[B][f]); CMP(fb[B][f]); CMP(fpos[B][f]); CMP(fneg[B][f]);
        for (int k = 0; k < a.flen[B][f]; ++k) { CMP(fv[B][a.foff[B][f] + k]); CMP(fe[B][a.foff[B][f] + k]); }
    }
    for (int c = 0; c < a.nc; ++c) CMP(cmat[c]);
    for (int p = 0; p < a.nm; ++p) CMP(umi[p]);
#undef CMP
    return 0;
}

template <class Caps>
struct Shared
{
    std::vector<Case> cases;
    MIComplex<Caps> cx;
    MIWarpScratch<Caps> sc;
    std::vector<MIComplex<Caps>> results;
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
        if (lane == 0) S.cx.init(C.v[0]);
        __syncwarp();
        for (int j = 1; j < C.k; ++j) warp_insert_material(S.cx, S.sc, C.v[j], lane);
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
        C.k = 2 + (int)(rng() % (maxk - 1));
        const int flavour = (int)(rng() % 8);
        for (int j = 0; j < C.k; ++j) {
            for (int c = 0; c < 4; ++c) {
                double x = U(rng);
                if (flavour == 1) x = (double)((int)(rng() % 4));              // small integers: ties at corners
                if (flavour == 2 && j > 0 && (rng() % 3) == 0) x = C.v[0][c]; // equal to material 0 at a corner
                if (flavour == 3) x = std::round(x * 4.0) / 4.0;              // coarse lattice
                C.v[j][c] = x;
            }
            if (flavour == 4 && j > 0 && (rng() % 2) == 0)                     // duplicate materials
                for (int c = 0; c < 4; ++c) C.v[j][c] = C.v[j - 1][c];
            if (flavour == 5 && j > 0)                                        // nearly identical materials
                for (int c = 0; c < 4; ++c) C.v[j][c] = C.v[0][c] + (c == j % 4 ? 1e-15 : 0.0);
            if (flavour == 6 && j > 0) {                                      // ties on a whole tet face
                for (int c = 0; c < 3; ++c) C.v[j][c] = C.v[0][c];
            }
        }
        S->cases.push_back(C);
    }
    S->results.resize(n_cases);
    std::memset(&simt::g_warp, 0, sizeof(simt::g_warp));
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
    long exact = 0, n_dup = 0, n_multi = 0, sum_cells = 0, n_bsplit = 0;
    for (int ci = 0; ci < n_cases && bad < 5; ++ci) {
        const Case& C = S->cases[ci];
        MIComplex<Caps>* ser = new MIComplex<Caps>();
        ser->init(C.v[0]);
        for (int j = 1; j < C.k; ++j) ser->insert(C.v[j]);
        n_dup += ser->has_dup;
        n_multi += ser->nc > 1;
        sum_cells += ser->nc;
        n_bsplit += ser->nf > 4 + (ser->nc > 1);
        const MIComplex<Caps>& wr = S->results[ci];
        n_err[ser->err]++;
        exact += ser->n_exact;
        if (ser->err == 1 && wr.err == 0) { // the serial capacity check is conservative
            ++cap_diff;
            delete ser;
            continue;
        }
        if (compare(*ser, wr, name, ci)) ++bad;
        delete ser;
    }
    std::printf("%s: %d cases, ok=%d degenerate=%d capacity=%d (serial-only capacity %d), exact fallbacks %ld, mismatches %d\n",
        name, n_cases, n_err[0], n_err[2], n_err[1], cap_diff, exact, bad);
    std::printf("   coverage: %ld with duplicate materials, %ld with > 1 cell (mean %.2f cells), %ld with split boundary faces\n",
        n_dup, n_multi, (double)sum_cells / n_cases, n_bsplit);
    delete S;
    return bad;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? std::atoi(argv[1]) : 2000;
    const unsigned seed = argc > 2 ? (unsigned)std::atoi(argv[2]) : 1u;
    int bad = 0;
    bad += run<MICapsSmall>(n, seed, 7, "small tier caps (k<=7: more than the tier holds -> capacity path)");
    bad += run<MICaps>(n / 4 + 1, seed + 7, 9, "big tier caps (k<=9)");
    return bad ? 1 : 0;
}
