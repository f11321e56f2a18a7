"""Golden digests of the cell-grouping maps (second extract_iso_mesh / extract_MI_mesh overloads,
/root/reference/src/extract_mesh.cpp:268-566, :988-1443), produced by the REFERENCE'S OWN functions through
oracle/_ref/libref_hybrid.so (needs /root/reference to build it).  Writes tests/golden/cellgroup_golden.json.

    python tests/golden/make_cellgroup_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import (MAP_I64, make_funcs, orc_eval, orc_grid, orc_run, ref_cellgroup_maps,  # noqa: E402
                     ref_mi_cellgroup_maps, synthetic_functions)

DEGENERATE_IA = [{"type": "plane", "point": [0.0, 0.0, 0.0], "normal": [1.0, 0.0, 0.0]},
                 {"type": "sphere", "center": [0.1, 0.05, -0.02], "radius": 0.6, "squared": True}]
DEGENERATE_MI = [{"type": "sphere", "center": [-0.5, 0, 0], "radius": 0.7},
                 {"type": "sphere", "center": [0.5, 0, 0], "radius": 0.7},
                 {"type": "sphere", "center": [0, 0.5, 0], "radius": 0.6},
                 {"type": "sphere", "center": [0, -0.5, 0], "radius": 0.6}]
CASES = {
    "ia_C2_R12": ("ia", 12, synthetic_functions("C2")),
    "ia_C4_R24": ("ia", 24, synthetic_functions("C4")),
    "ia_plane_through_vertices_R8": ("ia", 8, DEGENERATE_IA),
    "mi_C3_R16": ("mi", 16, synthetic_functions("C3")),
    "mi_ties_on_tet_faces_R8": ("mi", 8, DEGENERATE_MI),
    "mi_duplicate_material_R10": ("mi", 10, synthetic_functions("C3")[:4] + [dict(synthetic_functions("C3")[1])]),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.int64).tobytes()).hexdigest()


def reference_maps(mode, R, specs):
    pts, tets = orc_grid(R)
    vals = orc_eval(make_funcs(specs), pts)
    if mode == "ia":
        return ref_cellgroup_maps(tets, vals, len(pts))
    port = orc_run("mi", pts, tets, vals)  # the active-material lists are an input of the reference function
    return ref_mi_cellgroup_maps(tets, vals, port["func_in_tet"], port["start_index_of_tet"])


def main():
    out = {}
    for name, (mode, R, specs) in CASES.items():
        ref = reference_maps(mode, R, specs)
        assert ref.error == "", ref.error
        out[name] = {"counts": ref["counts"].tolist(), **{k: sha(ref[k]) for k in MAP_I64},
                     "n_entries": [int(len(ref[k])) for k in MAP_I64]}
    with open(os.path.join(HERE, "cellgroup_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
