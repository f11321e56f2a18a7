"""GPU parity tests of the material-interface path (src/material_interface.cpp hot path) through
the C-ABI: against the CPU oracle, the reference's golden inputs and BASELINE config C3."""
import hashlib
import json
import os

import numpy as np
import pytest

from gpu_compare import compare_mi
from helpers import (FLAG_LOOKUP, FLAG_SECONDARY, load_funcs, make_funcs, orc_eval, orc_grid, orc_run,
                     splitmix64, synthetic_functions)

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def rin():
    import rin_b200
    return rin_b200


@pytest.fixture(scope="module")
def ctx(rin):
    c = rin.Context(0)
    yield c
    c.close()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gpu_digest(mesh):
    return {"face_offsets": sha(mesh["face_offsets"].astype(np.int64)),
            "face_verts": sha(mesh["face_verts"].astype(np.int64)),
            "face_tets": sha(mesh["face_tets"].astype(np.int64).ravel()),
            "face_funcs_first": sha(mesh["face_funcs"].astype(np.int64)[:, 0]),
            "face_funcs": sha(mesh["face_funcs"].astype(np.int64).ravel()),
            "vert_xyz": sha(mesh["vert_xyz"])}


@pytest.mark.parametrize("R", [6, 21, 48])
@pytest.mark.parametrize("flags", [0, FLAG_LOOKUP, FLAG_LOOKUP | FLAG_SECONDARY])
def test_c3_six_spheres(ctx, rin, R, flags):
    """BASELINE C3 function set (6 overlapping spheres, triple junctions) on generated grids."""
    funcs = make_funcs(synthetic_functions("C3"))
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    port = orc_run("mi", pts, tets, vals, flags=flags)
    assert port.error == ""
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    cnt = ctx.run(rin.MODE_MI, flags)
    if flags == (FLAG_LOOKUP | FLAG_SECONDARY):
        # the 3-material ("secondary") table serves the 3-material tets: only >= 4 materials are left over
        assert cnt.num_k2 > 0 and cnt.num_general_tets <= cnt.num_kmore + cnt.num_k2 // 20 + cnt.num_k1 // 50
    elif flags == FLAG_LOOKUP:
        assert cnt.num_general_tets >= cnt.num_k2  # 3-material tets take the general kernel (:320-324)
    compare_mi(ctx, ctx.download_mesh(), port, cnt)
    if R == 48:
        assert cnt.num_k2 > 0  # three-material tets exist


def test_many_materials(ctx, rin):
    g = splitmix64(8)
    specs = [{"type": "sphere", "center": [1.6 * next(g) - 0.8 for _ in range(3)], "radius": 0.2 + 0.5 * next(g)}
             for _ in range(20)]
    funcs = make_funcs(specs)
    pts, tets = orc_grid(16)
    vals = orc_eval(funcs, pts)
    port = orc_run("mi", pts, tets, vals)
    ctx.set_mesh(pts, tets)
    ctx.set_values(vals)
    cnt = ctx.run(rin.MODE_MI, FLAG_LOOKUP | FLAG_SECONDARY)
    assert cnt.num_kmore > 0
    compare_mi(ctx, ctx.download_mesh(), port, cnt)


with open(os.path.join(G, "mi_goldens.json")) as _f:
    MI_GOLD = json.load(_f)


@pytest.mark.parametrize("name", sorted(k for k in MI_GOLD if k != "8-sphere"))
def test_reference_mi_golden_cases(ctx, rin, name):
    """tests/test_implicit_networks.cpp:475-683 inputs: counts + digests recorded from the
    reference's own extract_MI_mesh / compute_MI_vert_xyz."""
    funcs = load_funcs(os.path.join(G, "functions", name + ".json"))
    ctx.generate_grid(101)
    ctx.set_functions(funcs)
    cnt = ctx.run(rin.MODE_MI, FLAG_LOOKUP | FLAG_SECONDARY)
    st = MI_GOLD[name]["stats"]
    assert [cnt.num_intersecting_tet, cnt.num_k1, cnt.num_k2, cnt.num_kmore, cnt.num_verts, cnt.num_faces] == \
        [st[k] for k in ("num_intersecting_tet", "num_2_func", "num_3_func", "num_more_func", "num_MI_verts",
                         "num_MI_faces")]
    d = gpu_digest(ctx.download_mesh())
    for k, v in MI_GOLD[name]["digest"].items():
        assert d[k] == v, k


def test_reference_mi_eight_spheres_unstructured(ctx, rin):
    """tests/test_implicit_networks.cpp:685-731: 8 spheres on examples/tests/mesh.json."""
    d = np.load(os.path.join(G, "mi_8sphere_inputs.npz"))
    funcs = load_funcs(os.path.join(G, "functions", "8-sphere.json"))
    ctx.set_mesh(d["pts"], d["tets"])
    ctx.set_functions(funcs)
    cnt = ctx.run(rin.MODE_MI, FLAG_LOOKUP | FLAG_SECONDARY)
    mesh = ctx.download_mesh()
    gd = gpu_digest(mesh)
    for k, v in MI_GOLD["8-sphere"]["digest"].items():
        assert gd[k] == v, k
    vals = orc_eval(funcs, d["pts"])
    compare_mi(ctx, mesh, orc_run("mi", d["pts"], d["tets"], vals), cnt)


DEGENERATE = {
    # x and -x tie on the plane x = 0 through grid vertices: the whole interface lies on tet faces
    "x_vs_negx": [{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                  {"type": "plane", "point": [0, 0, 0], "normal": [-1, 0, 0]}],
    # spheres mirrored about grid planes: parts of the interface lie on tet faces
    "sym_spheres": [{"type": "sphere", "center": [-0.5, 0, 0], "radius": 0.7},
                    {"type": "sphere", "center": [0.5, 0, 0], "radius": 0.7},
                    {"type": "sphere", "center": [0, 0.5, 0], "radius": 0.6},
                    {"type": "sphere", "center": [0, -0.5, 0], "radius": 0.6}],
}


@pytest.mark.parametrize("name", sorted(DEGENERATE))
@pytest.mark.parametrize("R", [8, 20])
def test_materials_tying_on_tet_faces(ctx, rin, name, R):
    """Degenerate boundary-face matching of extract_MI_mesh (src/extract_mesh.cpp:833-981): the
    second tet that sees a tet face whose two sides hold different materials emits it."""
    pts, tets = orc_grid(R)
    funcs = make_funcs(DEGENERATE[name])
    vals = orc_eval(funcs, pts)
    port = orc_run("mi", pts, tets, vals)
    assert port.error == ""
    ctx.set_mesh(pts, tets)
    ctx.set_functions(funcs)
    cnt = ctx.run(rin.MODE_MI, FLAG_LOOKUP | FLAG_SECONDARY)
    compare_mi(ctx, ctx.download_mesh(), port, cnt)
