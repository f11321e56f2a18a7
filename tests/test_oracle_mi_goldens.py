"""CPU: the port oracle's material-interface path against the committed golden fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import load_funcs, orc_eval, orc_grid, orc_run

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(G, "mi_goldens.json")) as _f:
    MI_GOLD = json.load(_f)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(b):
    return {"face_offsets": sha(b["face_offsets"]), "face_verts": sha(b["face_verts"]),
            "face_tets": sha(b["face_tets"]), "face_funcs_first": sha(b["face_funcs"][0::2]),
            "face_funcs": sha(b["face_funcs"]), "vert_xyz": sha(b["vert_xyz"])}


@pytest.fixture(scope="module")
def grid101():
    return orc_grid(101)


@pytest.mark.parametrize("name", sorted(k for k in MI_GOLD if k != "8-sphere"))
def test_port_mi_golden(name, grid101):
    pts, tets = grid101
    vals = orc_eval(load_funcs(os.path.join(G, "functions", name + ".json")), pts)
    b = orc_run("mi", pts, tets, vals)
    assert b.error == ""
    st = MI_GOLD[name]["stats"]
    assert b["stats"].tolist()[3:] == [st[k] for k in ("num_intersecting_tet", "num_2_func", "num_3_func",
                                                       "num_more_func", "num_MI_verts", "num_MI_faces")]
    assert digest(b) == MI_GOLD[name]["digest"]


def test_port_mi_eight_spheres():
    d = np.load(os.path.join(G, "mi_8sphere_inputs.npz"))
    vals = orc_eval(load_funcs(os.path.join(G, "functions", "8-sphere.json")), d["pts"])
    b = orc_run("mi", d["pts"], d["tets"], vals)
    assert digest(b) == MI_GOLD["8-sphere"]["digest"]
