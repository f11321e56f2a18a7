"""BASELINE configurations at their FULL sizes on the GPU against committed oracle digests
(tests/golden/fullsize_golden.json, written by tests/golden/make_fullsize_golden.py from oracle/port and, where
it is fast enough, from the reference's own sources compiled in place), plus the cylinder / torus evaluator
against the oracle directly.  Per-tet combinatorics are pinned in aggregate only (the per-tet library is
un-vendored upstream): what is compared here is the extracted mesh, the active sets and
the vertex records, bit for bit."""
import hashlib
import json
import os

import numpy as np
import pytest

from gpu_compare import compare_ia
from helpers import make_funcs, orc_eval, orc_grid, orc_run, synthetic_functions

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(G, "fullsize_golden.json")) as _f:
    GOLD = json.load(_f)


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


def vert_rec(mesh, mi):
    """The oracle's vert_rec layout: tet, local, simplex size, 4 simplex vertices, 3 (IA) / 4 (MI) function ids."""
    sv = mesh["vert_simplex"].astype(np.int64)
    sv[sv == 0xFFFFFFFF] = -1
    fi = mesh["vert_funcs"].astype(np.int64)[:, :(4 if mi else 3)]
    fi[fi == 0xFFFFFFFF] = -1
    return np.concatenate([mesh["vert_tet"].astype(np.int64)[:, None], mesh["vert_local"].astype(np.int64)[:, None],
                           mesh["vert_simplex_size"].astype(np.int64)[:, None], sv, fi], axis=1)


def check_against_golden(ctx, rin, name):
    g = GOLD[name]
    mi = g["mode"] == "mi"
    ctx.generate_grid(g["grid"])
    ctx.set_functions(make_funcs(synthetic_functions(g["functions"])))
    cnt = ctx.run(rin.MODE_MI if mi else rin.MODE_IA)
    got = [cnt.num_pts, cnt.num_tets, cnt.num_degenerate_vertex, cnt.num_intersecting_tet, cnt.num_k1, cnt.num_k2,
           cnt.num_kmore, cnt.num_verts, cnt.num_faces]
    assert got == g["stats"], (got, g["stats"])
    mesh = ctx.download_mesh()
    ff = mesh["face_funcs"].astype(np.int64)
    ff[ff == 0xFFFFFFFF] = -1
    d = {"face_offsets": sha(mesh["face_offsets"].astype(np.int64)),
         "face_verts": sha(mesh["face_verts"].astype(np.int64)),
         "face_tets": sha(mesh["face_tets"].astype(np.int64).ravel()),
         "face_funcs_first": sha(ff[:, 0]), "vert_xyz": sha(mesh["vert_xyz"])}
    if mi:
        d["face_funcs"] = sha(ff.ravel())
    assert d == g["digest"]
    assert sha(vert_rec(mesh, mi)) == g["vert_rec"]
    fit, start = ctx.download_active()
    assert sha(fit.astype(np.int64)) == g["func_in_tet"]
    assert sha(start.astype(np.int64)) == g["start_index_of_tet"]
    return cnt


def test_c3_material_interface_128(ctx, rin):
    """BASELINE C3: material interface, 128^3, 6 materials (digest pinned by the reference's own sources)."""
    cnt = check_against_golden(ctx, rin, "C3")
    assert cnt.num_k2 > 0  # MI counters: k1 slot = 2 materials, k2 slot = 3 materials (secondary table)


def test_c4_dense_functions_128(ctx, rin):
    """BASELINE C4: 128^3, 32 near-coincident spheres, the general kernel dominates."""
    cnt = check_against_golden(ctx, rin, "C4")
    assert cnt.num_kmore > cnt.num_k1


def test_c5_256(ctx, rin):
    """BASELINE C5 on one GPU: 256^3 (83.9 M tets), 8 functions."""
    check_against_golden(ctx, rin, "C5")


def test_c2_with_cylinders_128(ctx, rin):
    """SURVEY 8(d) cylinder swap at 128^3 (digest pinned by the reference's own sources)."""
    check_against_golden(ctx, rin, "C2cyl")


def test_torus_cylinder_set_64(ctx, rin):
    check_against_golden(ctx, rin, "TOR")


@pytest.mark.parametrize("fset,R", [("TOR", 24), ("C2cyl", 30)])
def test_cylinder_torus_values_and_mesh_equal_the_oracle(ctx, fset, R):
    """Every parametric type of the device evaluator: values bit-exact, then the whole hot path element-wise."""
    funcs = make_funcs(synthetic_functions(fset))
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    assert np.array_equal(ctx.download_values(), vals)
    compare_ia(ctx, ctx.download_mesh(), orc_run("ia", pts, tets, vals), cnt)
    # the same through host points (unstructured path of the evaluator)
    ctx.set_mesh(pts, tets)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    assert np.array_equal(ctx.download_values(), vals)
    compare_ia(ctx, ctx.download_mesh(), orc_run("ia", pts, tets, vals), cnt)
