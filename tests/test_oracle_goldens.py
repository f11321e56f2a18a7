"""CPU: the port oracle against the committed golden fixtures (tests/golden, produced by
make_golden.py from the reference's own sources + fixture files) and, where the hybrid reference
library is present, against the reference's code directly."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import (FLAG_LOOKUP, FLAG_SECONDARY, load_funcs, make_funcs, orc_eval, orc_grid, orc_run,
                     ref_lib, ref_run, synthetic_functions)

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(bag):
    return {"face_offsets": sha(bag["face_offsets"]), "face_verts": sha(bag["face_verts"]),
            "face_tets": sha(bag["face_tets"]), "face_funcs_first": sha(bag["face_funcs"][0::2]),
            "vert_xyz": sha(bag["vert_xyz"])}


@pytest.fixture(scope="module")
def grid101():
    return orc_grid(101)


with open(os.path.join(G, "ia_goldens.json")) as _f:
    IA_GOLD = json.load(_f)


@pytest.mark.parametrize("name", sorted(IA_GOLD))
def test_port_reproduces_reference_golden_hot_path(name, grid101):
    """Hot-path outputs (counts, face arrays, coordinates) on the reference's own test inputs
    (generate_tet_mesh(101), examples/tests/*.json) equal what the reference's code produced."""
    pts, tets = grid101
    vals = orc_eval(load_funcs(os.path.join(G, "functions", name + ".json")), pts)
    if name == "3-sphere-5":
        vals[:, 2] = -vals[:, 2]  # tests/test_implicit_networks.cpp:402-404
    b = orc_run("ia", pts, tets, vals)
    assert b.error == ""
    gold = IA_GOLD[name]
    st = gold["stats"]
    want = [st["num_pts"], st["num_tets"], st["num_degenerate_vertex"], st["num_intersecting_tet"],
            st["num_1_func"], st["num_2_func"], st["num_more_func"], st["num_iso_verts"], st["num_iso_faces"]]
    assert b["stats"].tolist() == want
    assert digest(b) == gold["digest"]


def test_port_reproduces_c1():
    """BASELINE config C1: examples/implicit_arrangement/config.json on tet5_grid_10k."""
    d = np.load(os.path.join(G, "c1_inputs.npz"))
    with open(os.path.join(G, "c1_golden.json")) as f:
        gold = json.load(f)
    vals = orc_eval(load_funcs(os.path.join(G, "functions", "18-sphere.json")), d["pts"])
    b = orc_run("ia", d["pts"], d["tets"], vals)
    st = gold["stats"]
    assert b["stats"].tolist()[2:] == [st["num_degenerate_vertex"], st["num_intersecting_tet"], st["num_1_func"],
                                       st["num_2_func"], st["num_more_func"], st["num_iso_verts"],
                                       st["num_iso_faces"]]
    assert digest(b) == gold["digest"]


def test_port_small_cases_elementwise():
    small = np.load(os.path.join(G, "small_cases.npz"))
    pts, tets = orc_grid(12)
    for name in ("2-planesphere", "3-sphere-3"):
        vals = orc_eval(load_funcs(os.path.join(G, "functions", name + ".json")), pts)
        b = orc_run("ia", pts, tets, vals)
        for k in ("face_offsets", "face_verts", "face_tets", "vert_xyz"):
            assert np.array_equal(b[k], small[name + "/" + k]), (name, k)
        assert np.array_equal(b["face_funcs"][0::2], small[name + "/face_funcs"][0::2])


def test_lookup_switches_do_not_change_the_result():
    """Fig. 16 style differential test: tables on / secondary off / tables off."""
    pts, tets = orc_grid(20)
    vals = orc_eval(make_funcs(synthetic_functions("C2")), pts)
    ref = orc_run("ia", pts, tets, vals, flags=FLAG_LOOKUP | FLAG_SECONDARY)
    for flags in (FLAG_LOOKUP, 0):
        b = orc_run("ia", pts, tets, vals, flags=flags)
        for k in ("stats", "face_offsets", "face_verts", "face_tets", "vert_rec", "vert_xyz", "func_in_tet"):
            assert np.array_equal(b[k], ref[k]), (flags, k)


@pytest.mark.skipif(ref_lib() is None, reason="hybrid reference not built (needs /root/reference)")
@pytest.mark.parametrize("cfg,R", [("C2", 24), ("C4", 16), ("degenerate", 20)])
def test_port_equals_reference_code(cfg, R):
    pts, tets = orc_grid(R)
    if cfg == "degenerate":  # plane x = 0 through grid vertices + sphere
        funcs = make_funcs([{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                            {"type": "sphere", "center": [0, 0, 0], "radius": 0.5, "squared": True}])
    else:
        funcs = make_funcs(synthetic_functions(cfg))
    vals = orc_eval(funcs, pts)
    p = orc_run("ia", pts, tets, vals)
    r = ref_run("ia", pts, tets, vals)
    assert p.error == ""
    for k in ("face_offsets", "face_verts", "face_tet_offsets", "face_tets", "vert_xyz"):
        assert np.array_equal(p[k], r[k]), k
    if cfg != "degenerate":  # the reference's func_index of boundary-coplanar faces is an OOB read
        assert np.array_equal(p["face_funcs"][0::2], r["face_funcs"][0::2])
    assert p["stats"].tolist()[2:] == [r.stats[k] for k in ("num_degenerate_vertex", "num_intersecting_tet",
                                                            "num_1_func", "num_2_func", "num_more_func",
                                                            "num_iso_verts", "num_iso_faces")]


# tests/test_implicit_networks.cpp:853-1033: (functions, expression, patches, chains, corners, patch_sign_label)
CSG_GOLD = {
    "sphere_and_not_sphere": ("1-sphere", 0, 0, 0, 0, []),
    "three_spheres_2": ("3-sphere-2", 1, 4, 2, 0, [1, 1, 1, 1]),
    "three_spheres_3": ("3-sphere-3", 2, 5, 5, 2, [1, 1, 0, 0, 1]),
    "plane_two_spheres": ("3-planesphere", 3, 4, 2, 0, [1, 1, 1, 1]),
}


@pytest.mark.skipif(ref_lib() is None, reason="hybrid reference not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(CSG_GOLD))
def test_reference_csg_known_answers_on_the_restated_engine(name, grid101):
    """The reference's csg() + host stages on the restated per-tet engine reproduce its CSG tests."""
    from helpers import crs, ref_csg
    fn, expr, npatch, nchain, ncorner, sign = CSG_GOLD[name]
    pts, tets = grid101
    vals = orc_eval(load_funcs(os.path.join(G, "functions", fn + ".json")), pts)
    b = ref_csg(pts, tets, vals, expr)
    assert b.error == "" and b["success"][0] == 1
    assert (len(crs(b, "patches")), len(crs(b, "chains"))) == (npatch, nchain)
    assert sum(1 for l in crs(b, "non_manifold_edges_of_vert") if len(l) > 2) == ncorner
    assert b["patch_sign_label"].tolist() == sign
