"""Sizing aid for the shared-memory tiers: dumps the per-tet plane values of a workload's general tets (>= 3
active functions) for scripts/size_caps.cpp, which reports the peak transient entity counts by function count.

    python scripts/size_caps.py C4 64 /tmp/c4_tets.bin
    g++ -std=c++17 -O2 -ffp-contract=off -o /tmp/size_caps scripts/size_caps.cpp && /tmp/size_caps /tmp/c4_tets.bin
"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from helpers import make_funcs, orc_eval, orc_grid, synthetic_functions  # noqa: E402


def main(config="C4", R=64, out="/tmp/tets.bin"):
    pts, tets = orc_grid(int(R))
    vals = orc_eval(make_funcs(synthetic_functions(config)), pts)
    F = vals.shape[1]
    act = np.zeros((len(tets), F), bool)
    for f in range(F):
        s = np.sign(vals[:, f])[tets]
        act[:, f] = (s > 0).any(1) & (s < 0).any(1)
    k = act.sum(1)
    sel = np.nonzero(k >= 3)[0]
    print(len(sel), "general tets; histogram of k:", np.bincount(k[sel]).tolist())
    with open(out, "wb") as fo:
        for t in sel:
            fs = np.nonzero(act[t])[0]
            fo.write(struct.pack("i", len(fs)))
            fo.write(np.ascontiguousarray(vals[tets[t]][:, fs].T).tobytes())


if __name__ == "__main__":
    main(*sys.argv[1:])
