"""GPU: the reference's known-answer tests (tests/test_implicit_networks.cpp) replayed END TO END
through the drop-in: hot path on the GPU (librin_b200 via the C++ host layer), per-tet complexes
from rin_get_complexes, every host topology stage = the reference's own code
(oracle/_ref/libref_gpu_dropin.so, built where /root/reference is present and shipped prebuilt).
Expectations are the reference test's own numbers (transcribed in tests/golden/*.json)."""
import json
import os

import numpy as np
import pytest

from helpers import crs, dropin_lib, load_funcs, make_funcs, orc_eval, orc_grid, ref_csg, ref_run

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(dropin_lib() is None, reason="drop-in library not built (needs /root/reference)")]
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(G, "ia_goldens.json")) as _f:
    IA_GOLD = json.load(_f)
with open(os.path.join(G, "mi_goldens.json")) as _f:
    MI_GOLD = json.load(_f)


@pytest.fixture(scope="module")
def grid101():
    return orc_grid(101)


@pytest.mark.parametrize("name", sorted(IA_GOLD))
def test_ia_known_answers_through_the_gpu_drop_in(name, grid101):
    pts, tets = grid101
    vals = orc_eval(load_funcs(os.path.join(G, "functions", name + ".json")), pts)
    if name == "3-sphere-5":
        vals[:, 2] = -vals[:, 2]
    b = ref_run("ia", pts, tets, vals, lib=dropin_lib())
    assert b.error == "" and b["success"][0] == 1
    exp = IA_GOLD[name]["reference_test_expectation"]
    F = vals.shape[1]
    assert len(crs(b, "patches")) == exp["patches"]
    assert len(crs(b, "chains")) == exp["chains"]
    assert len(crs(b, "cells")) == exp["cells"]
    assert b["patch_function_label"].tolist() == exp["patch_function_label"]
    assert b["cell_function_label"].reshape(-1, F).tolist() == exp["cell_function_label"]
    for k in ("num_iso_verts", "num_iso_faces", "num_iso_edges", "num_patches", "num_chains", "num_shells",
              "num_components", "num_cells"):
        assert b.stats[k] == IA_GOLD[name]["stats"][k], k


@pytest.mark.parametrize("name", sorted(k for k in MI_GOLD if k != "8-sphere"))
def test_mi_known_answers_through_the_gpu_drop_in(name, grid101):
    pts, tets = grid101
    vals = orc_eval(load_funcs(os.path.join(G, "functions", name + ".json")), pts)
    b = ref_run("mi", pts, tets, vals, lib=dropin_lib())
    assert b.error == "" and b["success"][0] == 1
    exp = MI_GOLD[name]["reference_test_expectation"]
    assert len(crs(b, "patches")) == exp["patches"]
    assert len(crs(b, "chains")) == exp["chains"]
    assert len(crs(b, "cells")) == exp["cells"]
    assert b["patch_function_label"].reshape(-1, 2).tolist() == exp["patch_function_label"]
    assert b["cell_function_label"].tolist() == exp["cell_function_label"]


def test_mi_eight_spheres_through_the_gpu_drop_in():
    """tests/test_implicit_networks.cpp:685-731 on examples/tests/mesh.json: 8 shells, 8 cells, 6 corners and
    the 19-pair patch label vector; the symmetric spheres tie on tet faces (degenerate matching path)."""
    d = np.load(os.path.join(G, "mi_8sphere_inputs.npz"))
    vals = orc_eval(load_funcs(os.path.join(G, "functions", "8-sphere.json")), d["pts"])
    b = ref_run("mi", d["pts"], d["tets"], vals, lib=dropin_lib())
    assert b.error == "" and b["success"][0] == 1
    exp = MI_GOLD["8-sphere"]["reference_test_expectation"]
    assert len(crs(b, "shells")) == exp["shells"] and len(crs(b, "cells")) == exp["cells"]
    assert sum(1 for l in crs(b, "non_manifold_edges_of_vert") if len(l) > 2) == exp["corners"]
    assert b["patch_function_label"].reshape(-1, 2).tolist() == exp["patch_function_label"]
    assert b["cell_function_label"].tolist() == exp["cell_function_label"]


def test_c1_example_config_through_the_gpu_drop_in():
    """examples/implicit_arrangement/config.json (18 spheres on tet5_grid_10k): 1944 patches, 672 cells."""
    d = np.load(os.path.join(G, "c1_inputs.npz"))
    with open(os.path.join(G, "c1_golden.json")) as f:
        gold = json.load(f)
    vals = orc_eval(load_funcs(os.path.join(G, "functions", "18-sphere.json")), d["pts"])
    b = ref_run("ia", d["pts"], d["tets"], vals, lib=dropin_lib())
    assert b.error == "" and b["success"][0] == 1
    for k, v in gold["stats"].items():
        assert b.stats[k] == v, k
    assert b["patch_function_label"].tolist() == gold["patch_function_label"]


# tests/test_implicit_networks.cpp:853-1033: (functions, expression, patches, chains, corners, patch_sign_label)
CSG_GOLD = {
    "sphere_and_not_sphere": ("1-sphere", 0, 0, 0, 0, []),
    "three_spheres_2": ("3-sphere-2", 1, 4, 2, 0, [1, 1, 1, 1]),
    "three_spheres_3": ("3-sphere-3", 2, 5, 5, 2, [1, 1, 0, 0, 1]),
    "plane_two_spheres": ("3-planesphere", 3, 4, 2, 0, [1, 1, 1, 1]),
}


@pytest.mark.parametrize("name", sorted(CSG_GOLD))
def test_csg_known_answers_through_the_gpu_drop_in(name, grid101):
    """The reference's csg.cpp, linked unchanged, calls the GPU-backed implicit_arrangement."""
    fn, expr, npatch, nchain, ncorner, sign = CSG_GOLD[name]
    pts, tets = grid101
    vals = orc_eval(load_funcs(os.path.join(G, "functions", fn + ".json")), pts)
    b = ref_csg(pts, tets, vals, expr, lib=dropin_lib())
    assert b.error == "" and b["success"][0] == 1
    assert len(crs(b, "patches")) == npatch
    assert len(crs(b, "chains")) == nchain
    assert sum(1 for l in crs(b, "non_manifold_edges_of_vert") if len(l) > 2) == ncorner
    assert b["patch_sign_label"].tolist() == sign


def test_robust_test_mode_through_the_gpu_drop_in(grid101):
    """-R (src/implicit_arrangement.cpp:137-243, 329-343): returns true ("success.") and no mesh."""
    pts, tets = grid101
    vals = orc_eval(load_funcs(os.path.join(G, "functions", "3-sphere-3.json")), pts)
    b = ref_run("ia", pts, tets, vals, robust=True, lib=dropin_lib())
    assert b.error == "" and b["success"][0] == 1
    assert len(b["face_offsets"]) <= 1


def nested_components_case():
    """Four components (two disjoint spheres, a sphere nested inside one of them, a plane above): the cells cannot be
    found from the shells alone, the nesting order has to be resolved (src/implicit_arrangement.cpp:573-622)."""
    pts, tets = orc_grid(24)
    specs = [{"type": "sphere", "center": [-0.45, 0.0, 0.0], "radius": 0.3, "squared": True},
             {"type": "sphere", "center": [0.45, 0.1, 0.0], "radius": 0.32, "squared": True},
             {"type": "sphere", "center": [0.5, 0.1, 0.05], "radius": 0.12, "squared": True},
             {"type": "plane", "point": [0.0, 0.0, 0.6], "normal": [0.1, 0.0, 1.0]}]
    return pts, tets, orc_eval(make_funcs(specs), pts)


def test_cell_grouping_mode_through_the_gpu_drop_in():
    """useTopoRayShooting = false: maps of the second extract_iso_mesh overload from the device
    (rin_tet_maps), simplicial-cell graph from host/cell_graph.h; equal to the CPU reference in the same
    mode, and the same number of cells as ray shooting (the reference's Fig. 14 differential check)."""
    pts, tets, vals = nested_components_case()
    gpu = ref_run("ia", pts, tets, vals, ray=False, lib=dropin_lib())
    assert gpu.error == "" and gpu["success"][0] == 1
    cpu = ref_run("ia", pts, tets, vals, ray=False)
    assert gpu.stats["num_components"] == cpu.stats["num_components"] == 4
    # cells in the same order; the shells inside a cell come out of a hash set in the reference (unspecified order)
    assert [sorted(c) for c in crs(gpu, "cells")] == [sorted(c) for c in crs(cpu, "cells")]
    assert np.array_equal(gpu["cell_function_label"], cpu["cell_function_label"])
    shot = ref_run("ia", pts, tets, vals, ray=True, lib=dropin_lib())
    assert shot.stats["num_cells"] == gpu.stats["num_cells"] == 5
    assert sorted(map(sorted, crs(shot, "cells"))) == sorted(map(sorted, crs(gpu, "cells")))
    # timings.json keeps the reference's key set in both modes (src/implicit_arrangement.cpp:62-637)
    assert list(gpu.timing_labels) == list(cpu.timing_labels)
    assert list(shot.timing_labels) == list(ref_run("ia", pts, tets, vals, ray=True).timing_labels)
    assert len(gpu["timings"]) == len(gpu.timing_labels) and (np.asarray(gpu["timings"]) >= 0).all()


def test_mi_cell_grouping_mode_through_the_gpu_drop_in():
    """Material-interface analogue (src/material_interface.cpp:625-672): two components (a lone ball and two
    overlapping balls in a zero background material)."""
    pts, tets = orc_grid(20)
    specs = [{"type": "zero"},
             {"type": "sphere", "center": [-0.45, 0.03, 0.02], "radius": 0.3},
             {"type": "sphere", "center": [0.45, 0.1, 0.0], "radius": 0.32},
             {"type": "sphere", "center": [0.62, 0.1, 0.05], "radius": 0.3}]
    vals = orc_eval(make_funcs(specs), pts)
    gpu = ref_run("mi", pts, tets, vals, ray=False, lib=dropin_lib())
    assert gpu.error == "" and gpu["success"][0] == 1
    cpu = ref_run("mi", pts, tets, vals, ray=False)
    assert gpu.stats["num_components"] == cpu.stats["num_components"] == 2
    # cells in the same order; the shells inside a cell come out of a hash set in the reference (unspecified order)
    assert [sorted(c) for c in crs(gpu, "cells")] == [sorted(c) for c in crs(cpu, "cells")]
    assert np.array_equal(gpu["cell_function_label"], cpu["cell_function_label"])
    shot = ref_run("mi", pts, tets, vals, ray=True, lib=dropin_lib())
    assert shot.stats["num_cells"] == gpu.stats["num_cells"] == 4
    assert list(gpu.timing_labels) == list(cpu.timing_labels)
    assert list(shot.timing_labels) == list(ref_run("mi", pts, tets, vals, ray=True).timing_labels)
    assert len(gpu["timings"]) == len(gpu.timing_labels)


N1_CASES = [("ia", "C2", 24), ("ia", "3-sphere-3", 40), ("ia", "2-planesphere", 31), ("mi", "C3", 21),
            ("mi", "3-sphere-1", 36)]


@pytest.mark.parametrize("mode,fset,R", N1_CASES)
def test_edges_patches_chains_equal_the_reference_own_functions(mode, fset, R):
    """edges from the device (rin_mesh_edges), patches / chains from this repository's host layer,
    element-wise against compute_mesh_edges / compute_patches / compute_chains of the reference
    (src/mesh_connectivity.cpp:10-56, 58-94, 196-241) run by the hybrid on the same inputs: edge ids, end
    vertices, face_edge_indices, patch order and the order of faces inside a patch, chain order."""
    from helpers import synthetic_functions
    pts, tets = orc_grid(R)
    if fset == "degenerate":
        funcs = make_funcs([{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                            {"type": "sphere", "center": [0, 0, 0], "radius": 0.5, "squared": True},
                            {"type": "plane", "point": [0, 0, 0], "normal": [0, 1, 0]}])
    elif fset in ("C2", "C3"):
        funcs = make_funcs(synthetic_functions(fset))
    else:
        funcs = load_funcs(os.path.join(G, "functions", fset + ".json"))
    vals = orc_eval(funcs, pts)
    ref = ref_run(mode, pts, tets, vals)
    got = ref_run(mode, pts, tets, vals, lib=dropin_lib())
    # (the reference's MI face ordering gives up when an edge's faces span several tets, src/pair_faces.cpp:131-134:
    #  both sides then report the same failure, after the N1 stages have run)
    assert ref.error == "" and got.error == "" and got["success"][0] == ref["success"][0]
    for k in ("edges", "edge_faces", "edge_faces_offsets", "patches", "patches_offsets", "chains", "chains_offsets",
              "non_manifold_edges_of_vert", "non_manifold_edges_of_vert_offsets", "patch_function_label"):
        assert np.array_equal(got[k], ref[k]), k
    assert len(ref["edges"]) > 0


def test_complexes_are_fetched_lazily(grid101):
    """the host stages read the complexes of a few tets only (one per chain, plus the owners of
    iso-vertices on spanning-forest edges when there are several components); the drop-in fetches exactly those
    from the device instead of all active tets - and the reference's known answers still come out."""
    pts, tets = grid101
    for name, bound in (("2-planesphere", 0.01), ("3-sphere-3", 0.01), ("3-sphere-1", 0.35)):
        vals = orc_eval(load_funcs(os.path.join(G, "functions", name + ".json")), pts)
        b = ref_run("ia", pts, tets, vals, lib=dropin_lib())
        assert b.error == "" and b["success"][0] == 1
        exp = IA_GOLD[name]["reference_test_expectation"]
        assert len(crs(b, "cells")) == exp["cells"] and len(crs(b, "chains")) == exp["chains"]
        active = b.stats["num_intersecting_tet"]
        n = _fetched(b)
        assert n is not None and n <= bound * active + 64, (name, n, active)


def _fetched(bag):
    import ctypes as C
    n = C.c_uint64()
    p = bag._lib.ref_i64(bag._h, b"complexes_fetched", C.byref(n))
    return int(p[0]) if n.value else None
