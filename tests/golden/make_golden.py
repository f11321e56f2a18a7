"""Generates the committed golden fixtures.  Run HERE (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

Sources of truth:
 * the reference's own known-answer tests (/root/reference/tests/test_implicit_networks.cpp:56-435,
   475-731): patch / chain / cell counts and label vectors, transcribed into EXPECTED below;
 * the hybrid reference oracle/_ref/libref_hybrid.so (the reference's src/*.cpp compiled in place on
   top of the restated per-tet engine), run on the reference's fixture files
   (/root/reference/examples/tests/*.json, examples/tet_mesh/tet5_grid_10k.json,
   examples/implicit_arrangement/18-sphere.json).
Outputs (small): ia_goldens.json, c1_inputs.npz, small_cases.npz, functions/*.json (function
parameter files, copied as data so that the tests do not need /root/reference at run time).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import *  # noqa

REF_EX = "/root/reference/examples"

# transcribed from tests/test_implicit_networks.cpp (patches, chains, cells, patch labels, cell labels)
EXPECTED_IA = {
    "1-plane": (1, 0, 2, [0], [[1], [0]]),
    "1-sphere": (1, 0, 2, [0], [[1], [0]]),
    "2-planesphere": (4, 1, 4, [1, 0, 0, 1], [[0, 1], [0, 0], [1, 0], [1, 1]]),
    "2-sphere": (4, 1, 4, [0, 1, 1, 0], [[1, 0], [0, 0], [1, 1], [0, 1]]),
    "3-sphere-1": (3, 0, 4, [0, 2, 1], [[1, 0, 0], [0, 0, 0], [0, 0, 1], [0, 1, 0]]),
    "3-sphere-2": (7, 2, 6, [0, 2, 0, 2, 1, 1, 2],
                   [[1, 0, 0], [0, 0, 0], [1, 0, 1], [0, 0, 1], [0, 1, 1], [0, 1, 0]]),
    "3-sphere-3": (12, 6, 8, [0, 2, 0, 2, 1, 1, 2, 1, 0, 1, 2, 0],
                   [[1, 0, 0], [0, 0, 0], [0, 0, 1], [1, 0, 1], [1, 1, 0], [1, 1, 1], [0, 1, 1], [0, 1, 0]]),
    "3-sphere-4": (3, 0, 4, [2, 1, 0], [[0, 0, 1], [0, 0, 0], [0, 1, 1], [1, 1, 1]]),
    # "three spheres 5": third function negated by the test (:402-404)
    "3-sphere-5": (4, 1, 4, [2, 1, 1, 2], [[1, 0, 1], [1, 0, 0], [1, 1, 1], [1, 1, 0]]),
}


# material interface goldens (tests/test_implicit_networks.cpp:475-731)
EXPECTED_MI = {
    "1-sphere": (0, 0, 1, [], [0]),
    "2-planesphere": (1, 0, 2, [[1, 0]], [1, 0]),
    "2-sphere": (1, 0, 2, [[1, 0]], [1, 0]),
    "3-sphere-1": (3, 1, 3, [[2, 0], [2, 1], [1, 0]], [2, 0, 1]),
    "3-sphere-4": (0, 0, 1, [], [2]),
}
EXPECTED_MI_8SPHERE = {
    "shells": 8, "cells": 8, "corners": 6,
    "patch_function_label": [[3, 1], [7, 3], [7, 6], [3, 2], [6, 2], [6, 4], [5, 4], [5, 1], [7, 5], [4, 0], [2, 0],
                             [1, 0], [7, 1], [7, 4], [5, 0], [6, 0], [3, 0], [7, 2], [7, 0]],
    "cell_function_label": [3, 1, 7, 6, 2, 4, 5, 0],
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mesh_digest(bag):
    return {
        "face_offsets": sha(bag["face_offsets"]), "face_verts": sha(bag["face_verts"]),
        "face_tets": sha(bag["face_tets"]), "face_funcs_first": sha(bag["face_funcs"][0::2]),
        "vert_xyz": sha(bag["vert_xyz"]),
    }


def main():
    os.makedirs(os.path.join(HERE, "functions"), exist_ok=True)
    out = {}
    pts, tets = orc_grid(101)
    for name, (npatch, nchain, ncell, plabel, clabel) in EXPECTED_IA.items():
        with open(os.path.join(REF_EX, "tests", name + ".json")) as f:
            spec = json.load(f)
        with open(os.path.join(HERE, "functions", name + ".json"), "w") as f:
            json.dump(spec, f)
        vals = orc_eval(make_funcs(spec), pts)
        if name == "3-sphere-5":
            vals[:, 2] = -vals[:, 2]
        b = ref_run("ia", pts, tets, vals)
        F = vals.shape[1]
        got = (len(crs(b, "patches")), len(crs(b, "chains")), len(crs(b, "cells")),
               b["patch_function_label"].tolist(), b["cell_function_label"].reshape(-1, F).tolist())
        assert got == (npatch, nchain, ncell, plabel, clabel), (name, got)
        out[name] = {"grid": 101, "reference_test_expectation": {
            "patches": npatch, "chains": nchain, "cells": ncell, "patch_function_label": plabel,
            "cell_function_label": clabel}, "stats": b.stats, "digest": mesh_digest(b)}
        print(name, "golden reproduced", b.stats["num_iso_verts"], b.stats["num_iso_faces"])
    with open(os.path.join(HERE, "ia_goldens.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

    # BASELINE config C1: examples/implicit_arrangement/config.json on tet5_grid_10k
    with open(os.path.join(REF_EX, "tet_mesh", "tet5_grid_10k.json")) as f:
        m = json.load(f)
    c1_pts = np.asarray(m[0], np.float64)
    c1_tets = np.asarray(m[1], np.uint32)
    with open(os.path.join(REF_EX, "implicit_arrangement", "18-sphere.json")) as f:
        spec = json.load(f)
    with open(os.path.join(HERE, "functions", "18-sphere.json"), "w") as f:
        json.dump(spec, f)
    vals = orc_eval(make_funcs(spec), c1_pts)
    b = ref_run("ia", c1_pts, c1_tets, vals)
    assert b.error == "" and b["success"][0] == 1
    np.savez_compressed(os.path.join(HERE, "c1_inputs.npz"), pts=c1_pts, tets=c1_tets)
    c1 = {"stats": b.stats, "digest": mesh_digest(b),
          "patch_function_label": b["patch_function_label"].tolist()}
    with open(os.path.join(HERE, "c1_golden.json"), "w") as f:
        json.dump(c1, f, indent=1, sort_keys=True)
    print("C1", b.stats)

    # ---- material interface
    mi_out = {}
    for name, (npatch, nchain, ncell, plabel, clabel) in EXPECTED_MI.items():
        with open(os.path.join(REF_EX, "tests", name + ".json")) as f:
            spec = json.load(f)
        vals = orc_eval(make_funcs(spec), pts)
        b = ref_run("mi", pts, tets, vals)
        got = (len(crs(b, "patches")), len(crs(b, "chains")), len(crs(b, "cells")),
               b["patch_function_label"].reshape(-1, 2).tolist(), b["cell_function_label"].tolist())
        assert got == (npatch, nchain, ncell, plabel, clabel), (name, got)
        d = mesh_digest(b)
        d["face_funcs"] = sha(b["face_funcs"])
        mi_out[name] = {"grid": 101, "reference_test_expectation": {
            "patches": npatch, "chains": nchain, "cells": ncell, "patch_function_label": plabel,
            "cell_function_label": clabel}, "stats": b.stats, "digest": d}
        print("MI", name, "golden reproduced", b.stats["num_MI_verts"], b.stats["num_MI_faces"])
    with open(os.path.join(REF_EX, "tests", "mesh.json")) as f:
        m = json.load(f)
    m_pts = np.asarray(m[0], np.float64)
    m_tets = np.asarray(m[1], np.uint32)
    with open(os.path.join(REF_EX, "tests", "8-sphere.json")) as f:
        spec = json.load(f)
    with open(os.path.join(HERE, "functions", "8-sphere.json"), "w") as f:
        json.dump(spec, f)
    vals = orc_eval(make_funcs(spec), m_pts)
    b = ref_run("mi", m_pts, m_tets, vals)
    corners = sum(1 for l in crs(b, "non_manifold_edges_of_vert") if len(l) > 2)
    got = {"shells": len(crs(b, "shells")), "cells": len(crs(b, "cells")), "corners": corners,
           "patch_function_label": b["patch_function_label"].reshape(-1, 2).tolist(),
           "cell_function_label": b["cell_function_label"].tolist()}
    assert got == EXPECTED_MI_8SPHERE, got
    np.savez_compressed(os.path.join(HERE, "mi_8sphere_inputs.npz"), pts=m_pts, tets=m_tets)
    d = mesh_digest(b)
    d["face_funcs"] = sha(b["face_funcs"])
    mi_out["8-sphere"] = {"mesh": "examples/tests/mesh.json", "reference_test_expectation": EXPECTED_MI_8SPHERE,
                          "stats": b.stats, "digest": d}
    print("MI 8-sphere golden reproduced", b.stats)
    with open(os.path.join(HERE, "mi_goldens.json"), "w") as f:
        json.dump(mi_out, f, indent=1, sort_keys=True)

    # small cases with full arrays (reference extract + xyz code on the restated engine)
    small = {}
    spts, stets = orc_grid(12)
    for name in ("2-planesphere", "3-sphere-3"):
        with open(os.path.join(REF_EX, "tests", name + ".json")) as f:
            spec = json.load(f)
        vals = orc_eval(make_funcs(spec), spts)
        b = ref_run("ia", spts, stets, vals)
        for k in ("face_offsets", "face_verts", "face_tets", "face_funcs", "vert_xyz", "stats"):
            small[name + "/" + k] = b[k]
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **small)


if __name__ == "__main__":
    main()
