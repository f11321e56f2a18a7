// Sizing aid: peak entity counts (transient, before re-packing) of the per-tet complexes of a workload,
// by number of active functions.  Input: binary file of records {int32 k; double v[k][4]} (scripts/size_caps.py).
#include "../tests/simt/simt_emul.h"
static int g_peak[6];
#define RIN_TRACK_PEAK(nv, ne, nf, nc, nfe, ncf)                                   \
    do {                                                                           \
        const int cur_[6] = {nv, ne, nf, nc, nfe, ncf};                            \
        for (int i_ = 0; i_ < 6; ++i_) if (cur_[i_] > g_peak[i_]) g_peak[i_] = cur_[i_]; \
    } while (0)
#include "../robust-implicit-surface-networks_b200/csrc/ia_complex.cuh"
#include <map>
using namespace rin;
int main(int argc, char** argv)
{
    FILE* f = std::fopen(argv[1], "rb");
    std::map<int, std::vector<long>> hist; // k -> maxima + count
    auto* cx = new IAComplex<IACaps>();
    int k;
    long n = 0, nerr = 0;
    std::map<int, std::map<int, long>> ne_hist;
    while (std::fread(&k, 4, 1, f) == 1) {
        double v[64][4];
        if (std::fread(v, 32, k, f) != (size_t)k) break;
        for (int i = 0; i < 6; ++i) g_peak[i] = 0;
        cx->init();
        for (int j = 0; j < k; ++j) cx->insert(v[j]);
        if (cx->err) { ++nerr; continue; }
        auto& h = hist[k];
        if (h.empty()) h.assign(7, 0);
        for (int i = 0; i < 6; ++i) h[i] = std::max<long>(h[i], g_peak[i]);
        h[6]++;
        ne_hist[k][g_peak[1] / 32]++;
        ++n;
    }
    std::printf("tets %ld errors %ld\n k   count   nv   ne   nf   nc  nfe  ncf (peaks)\n", n, nerr);
    for (auto& [kk, h] : hist)
        std::printf("%2d %7ld %4ld %4ld %4ld %4ld %4ld %4ld\n", kk, h[6], h[0], h[1], h[2], h[3], h[4], h[5]);
    for (auto& [kk, m] : ne_hist) {
        std::printf("k=%d peak-ne/32 histogram:", kk);
        for (auto& [b, c] : m) std::printf(" %d:%ld", b * 32, c);
        std::printf("\n");
    }
    return 0;
}
