"""rin_tet_maps vs the oracle's restatement of the cell-grouping extraction (oracle/port/port.cpp; second
extract_iso_mesh / extract_MI_mesh overloads, src/extract_mesh.cpp:268-566 and :988-1443) and,
where the hybrid reference library is present, vs the reference's own functions called directly."""
import os

import numpy as np
import pytest

from helpers import (MAP_I64, load_funcs, make_funcs, orc_eval, orc_grid, orc_run, ref_cellgroup_maps, ref_lib,
                     synthetic_functions)

pytestmark = [pytest.mark.gpu]


@pytest.fixture()
def ctx():
    import rin_b200 as rin
    c = rin.Context(0)
    yield c
    c.close()


def expand(maps, T):
    """active-tet CRS -> the reference's T+1-entry start arrays (inactive tets have empty ranges, :329-330)."""
    out = {}
    for kind in ("vert", "face"):
        off = maps[kind + "_offsets"].astype(np.int64)
        sizes = np.zeros(T, np.int64)
        sizes[maps["active_tets"]] = off[1:] - off[:-1]
        start = np.zeros(T + 1, np.int64)
        np.cumsum(sizes, out=start[1:])
        out[kind + "_start"] = start
    ff = maps["face_ids"].astype(np.int64)
    ff[ff == 0xFFFFFFFF] = -1
    out["face_ids"] = ff
    out["vert_ids"] = maps["vert_ids"]
    return out


def compare_maps(got, bag):
    """got: expand()-ed device maps; bag: oracle port run or reference function result (same four arrays)."""
    assert bag.error == "", bag.error
    assert np.array_equal(got["vert_start"], bag["global_vId_start_index_of_tet"])
    assert np.array_equal(got["vert_ids"], bag["global_vId_of_tet_vert"])
    assert np.array_equal(got["face_start"], bag["iso_fId_start_index_of_tet"])
    assert np.array_equal(got["face_ids"], bag["iso_fId_of_tet_face"])


def check(ctx, pts, tets, vals, **run_kw):
    ctx.set_mesh(pts, tets)
    ctx.set_values(vals)
    cnt = ctx.run(**run_kw)
    got = expand(ctx.tet_maps(), len(tets))
    port = orc_run("ia", pts, tets, vals)
    assert port["stats"][-2:].tolist() == [cnt.num_verts, cnt.num_faces]
    compare_maps(got, port)
    if ref_lib() is not None:
        ref = ref_cellgroup_maps(tets, vals, len(pts))
        assert ref["counts"].tolist() == [cnt.num_verts, cnt.num_faces]
        compare_maps(got, ref)
    return cnt


@pytest.mark.parametrize("cfg,R", [("C2", 16), ("C4", 24)])
def test_maps_on_synthetic_configs(ctx, cfg, R):
    pts, tets = orc_grid(R)
    vals = orc_eval(make_funcs(synthetic_functions(cfg)), pts)
    cnt = check(ctx, pts, tets, vals)
    assert cnt.num_kmore > 0


def test_maps_without_lookup_tables(ctx):
    import rin_b200 as rin
    pts, tets = orc_grid(10)
    vals = orc_eval(make_funcs(synthetic_functions("C2")), pts)
    check(ctx, pts, tets, vals, mode=rin.MODE_IA, flags=0)


def test_maps_with_iso_faces_on_tet_boundaries(ctx):
    """A plane through grid vertices: iso-vertices on tet corners (encoded -(id)-1, :518) and iso-faces shared by
    two tets (deduplicated ids, :536-551)."""
    pts, tets = orc_grid(8)
    specs = [{"type": "plane", "point": [0.0, 0.0, 0.0], "normal": [1.0, 0.0, 0.0]},
             {"type": "sphere", "center": [0.1, 0.05, -0.02], "radius": 0.6, "squared": True}]
    vals = orc_eval(make_funcs(specs), pts)
    cnt = check(ctx, pts, tets, vals)
    assert cnt.num_degenerate_vertex > 0 and cnt.num_face_tets > cnt.num_faces


def test_maps_on_reference_fixture(ctx):
    funcs = load_funcs(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "functions", "3-sphere-1.json"))
    pts, tets = orc_grid(12)
    vals = orc_eval(funcs, pts)
    check(ctx, pts, tets, vals)


# ---- material interface (second extract_MI_mesh overload, src/extract_mesh.cpp:988-1443) ----------------------

def check_mi(ctx, pts, tets, vals):
    import rin_b200 as rin
    from helpers import orc_run, ref_mi_cellgroup_maps
    port = orc_run("mi", pts, tets, vals)
    assert port.error == ""
    ctx.set_mesh(pts, tets)
    ctx.set_values(vals)
    cnt = ctx.run(rin.MODE_MI)
    got = expand(ctx.tet_maps(), len(tets))
    assert port["stats"][-2:].tolist() == [cnt.num_verts, cnt.num_faces]
    compare_maps(got, port)
    if ref_lib() is not None:
        ref = ref_mi_cellgroup_maps(tets, vals, port["func_in_tet"], port["start_index_of_tet"])
        assert ref["counts"].tolist() == [cnt.num_verts, cnt.num_faces]
        compare_maps(got, ref)
    return cnt


def test_mi_maps_on_c3(ctx):
    pts, tets = orc_grid(16)
    vals = orc_eval(make_funcs(synthetic_functions("C3")), pts)
    cnt = check_mi(ctx, pts, tets, vals)
    assert cnt.num_k2 > 0  # three-material tets: general kernel records


@pytest.mark.parametrize("name", ["x_vs_negx", "sym_spheres"])
def test_mi_maps_with_materials_tying_on_tet_faces(ctx, name):
    """Boundary faces that are material interfaces carry the same face id in both tets (:1391-1392)."""
    from test_gpu_mi import DEGENERATE
    pts, tets = orc_grid(8)
    vals = orc_eval(make_funcs(DEGENERATE[name]), pts)
    check_mi(ctx, pts, tets, vals)


def test_mi_maps_with_duplicate_materials(ctx):
    pts, tets = orc_grid(10)
    specs = synthetic_functions("C3")[:4]
    specs.append(dict(specs[1]))  # an identical copy of a material
    vals = orc_eval(make_funcs(specs), pts)
    check_mi(ctx, pts, tets, vals)
