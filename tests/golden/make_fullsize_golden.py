"""Generates tests/golden/fullsize_golden.json: digests of the CPU oracle's output on the BASELINE
configurations at their FULL sizes (C3 material interface 128^3, C4 32 functions 128^3, C5 256^3) and on the
cylinder / torus function sets.  Run HERE (minutes of CPU, ~10 GB of RAM for C5):

    python tests/golden/make_fullsize_golden.py [C3 C4 C5 C2cyl TOR ...]

Every case is computed by oracle/port (liboracle.so).  Where the hybrid reference (the reference's own src/*.cpp
compiled in place, oracle/_ref/libref_hybrid.so) finishes in reasonable time it is run on the same inputs and its
mesh digest must equal the port's: the recorded digest is then the reference's own output ("pinned_by": "reference").
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import (make_funcs, orc_eval, orc_grid, orc_run, ref_lib, ref_run,  # noqa: E402
                     synthetic_functions)

CASES = {  # name -> (mode, function set, R, run the hybrid reference too)
    "C2cyl": ("ia", "C2cyl", 128, True),
    "TOR": ("ia", "TOR", 64, True),
    "C3": ("mi", "C3", 128, True),
    "C4": ("ia", "C4", 128, True),
    "C5": ("ia", "C5", 256, True),
    # weak-scaling instances of bench.py: R = round(128 N^(1/3)) for N = 1, 2, 4 (N = 8 is C5)
    "C2": ("ia", "C2", 128, True),
    "W2": ("ia", "C2", 161, False),
    "W4": ("ia", "C2", 203, False),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mesh_digest(bag, mi):
    d = {"face_offsets": sha(bag["face_offsets"]), "face_verts": sha(bag["face_verts"]),
         "face_tets": sha(bag["face_tets"]), "face_funcs_first": sha(bag["face_funcs"][0::2]),
         "vert_xyz": sha(bag["vert_xyz"])}
    if mi:
        d["face_funcs"] = sha(bag["face_funcs"])
    return d


def main():
    path = os.path.join(HERE, "fullsize_golden.json")
    out = {}
    if os.path.exists(path):
        with open(path) as f:
            out = json.load(f)
    for name in (sys.argv[1:] or list(CASES)):
        mode, fset, R, with_ref = CASES[name]
        t0 = time.time()
        funcs = make_funcs(synthetic_functions(fset))
        pts, tets = orc_grid(R)
        vals = orc_eval(funcs, pts)
        port = orc_run(mode, pts, tets, vals)
        assert port.error == "", port.error
        rec = {"mode": mode, "functions": fset, "grid": R, "stats": port["stats"].tolist(),
               "values": sha(vals), "digest": mesh_digest(port, mode == "mi"),
               "func_in_tet": sha(port["func_in_tet"]), "start_index_of_tet": sha(port["start_index_of_tet"]),
               "vert_rec": sha(port["vert_rec"]), "pinned_by": "port"}
        print(name, "port", rec["stats"], "%.1fs" % (time.time() - t0), flush=True)
        if with_ref and ref_lib() is not None:
            t0 = time.time()
            b = ref_run(mode, pts, tets, vals)
            assert b.error == "" and b["success"][0] == 1, b.error
            assert mesh_digest(b, mode == "mi") == rec["digest"], name
            rec["pinned_by"] = "reference"
            rec["reference_stats"] = b.stats
            print(name, "reference agrees", "%.1fs" % (time.time() - t0), flush=True)
        out[name] = rec
        del port
        with open(path, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
