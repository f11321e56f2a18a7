"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C-ABI against the
CPU oracle on the same inputs, against the committed golden fixtures, and through size-independent
properties at the BASELINE sizes."""
import hashlib
import json
import os

import numpy as np
import pytest

from gpu_compare import compare_ia
from helpers import (FLAG_LOOKUP, FLAG_SECONDARY, load_funcs, make_funcs, orc_eval, orc_grid, orc_run,
                     synthetic_functions)

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
    ff = mesh["face_funcs"].astype(np.int64)[:, 0]
    return {"face_offsets": sha(mesh["face_offsets"].astype(np.int64)),
            "face_verts": sha(mesh["face_verts"].astype(np.int64)),
            "face_tets": sha(mesh["face_tets"].astype(np.int64).ravel()),
            "face_funcs_first": sha(ff), "vert_xyz": sha(mesh["vert_xyz"])}


@pytest.mark.parametrize("cfg,R", [("C2", 8), ("C2", 33), ("C4", 20), ("C2", 64)])
def test_generated_grid_and_parametric_functions(ctx, rin, cfg, R):
    """Device-side grid generation + function evaluation + full hot path, bit-exact."""
    funcs = make_funcs(synthetic_functions(cfg))
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    port = orc_run("ia", pts, tets, vals)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    gp, gt = ctx.download_grid(len(pts), len(tets))
    assert np.array_equal(gp, pts) and np.array_equal(gt.astype(np.uint64), tets)
    assert np.array_equal(ctx.download_values(), vals)  # function values: bit-exact
    compare_ia(ctx, ctx.download_mesh(), port, cnt)


@pytest.mark.parametrize("flags", [0, FLAG_LOOKUP, FLAG_LOOKUP | FLAG_SECONDARY])
def test_host_arrays_and_lookup_switches(ctx, flags):
    """The legacy signature path (host pts / size_t tets / row-major values) with every
    use_lookup / use_secondary_lookup combination (src/implicit_arrangement.cpp:276-284)."""
    pts, tets = orc_grid(18)
    vals = orc_eval(make_funcs(synthetic_functions("C2")), pts)
    port = orc_run("ia", pts, tets, vals, flags=flags)
    ctx.set_mesh(pts, tets)  # uint64 indices, narrowed on the device
    ctx.set_values(vals)
    cnt = ctx.run(flags=flags)
    compare_ia(ctx, ctx.download_mesh(), port, cnt)
    if flags == 0:
        assert cnt.num_general_tets == cnt.num_intersecting_tet


with open(os.path.join(G, "ia_goldens.json")) as _f:
    IA_GOLD = json.load(_f)


@pytest.mark.parametrize("name", sorted(IA_GOLD))
def test_reference_golden_cases(ctx, name):
    """The reference's known-answer inputs (tests/test_implicit_networks.cpp: grid 101, fixture
    functions): counts and array digests recorded from the reference's own code."""
    funcs = load_funcs(os.path.join(G, "functions", name + ".json"))
    ctx.generate_grid(101)
    if name == "3-sphere-5":
        funcs[2]["flip"] = 1  # the test negates the third column (:402-404)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    st = IA_GOLD[name]["stats"]
    assert [cnt.num_degenerate_vertex, cnt.num_intersecting_tet, cnt.num_k1, cnt.num_k2, cnt.num_kmore,
            cnt.num_verts, cnt.num_faces] == [st[k] for k in (
                "num_degenerate_vertex", "num_intersecting_tet", "num_1_func", "num_2_func", "num_more_func",
                "num_iso_verts", "num_iso_faces")]
    assert gpu_digest(ctx.download_mesh()) == IA_GOLD[name]["digest"]


def test_c1_unstructured_mesh(ctx):
    """BASELINE C1: examples/implicit_arrangement/config.json on tet5_grid_10k (18 spheres)."""
    d = np.load(os.path.join(G, "c1_inputs.npz"))
    with open(os.path.join(G, "c1_golden.json")) as f:
        gold = json.load(f)
    funcs = load_funcs(os.path.join(G, "functions", "18-sphere.json"))
    ctx.set_mesh(d["pts"], d["tets"])
    ctx.set_functions(funcs)
    cnt = ctx.run()
    assert cnt.num_verts == gold["stats"]["num_iso_verts"] and cnt.num_faces == gold["stats"]["num_iso_faces"]
    mesh = ctx.download_mesh()
    assert gpu_digest(mesh) == gold["digest"]
    vals = orc_eval(funcs, d["pts"])
    compare_ia(ctx, mesh, orc_run("ia", d["pts"], d["tets"], vals), cnt)


def test_degenerate_plane_through_grid_vertices(ctx):
    """Plane x = 0 on an even grid: zero signs, iso-vertices on tet vertices, iso-faces on tet
    faces shared by two tets (src/extract_mesh.cpp:197-225, 240-253)."""
    pts, tets = orc_grid(20)
    funcs = make_funcs([{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                        {"type": "sphere", "center": [0, 0, 0], "radius": 0.5, "squared": True},
                        {"type": "plane", "point": [0, 0, 0], "normal": [0, 1, 0]}])
    vals = orc_eval(funcs, pts)
    port = orc_run("ia", pts, tets, vals)
    assert port["stats"][2] > 0
    ctx.set_mesh(pts, tets.astype(np.uint32))
    ctx.set_functions(funcs)
    cnt = ctx.run()
    mesh = ctx.download_mesh()
    assert cnt.num_face_tets > cnt.num_faces  # some faces are shared by two tets
    compare_ia(ctx, mesh, port, cnt)


def test_many_functions_multiword_masks(ctx):
    """F = 40 > 32: two mask words per vertex."""
    from helpers import splitmix64
    pts, tets = orc_grid(20)
    g = splitmix64(6)
    specs = [{"type": "sphere", "center": [1.6 * next(g) - 0.8 for _ in range(3)],
              "radius": 0.1 + 0.4 * next(g), "squared": bool(i % 2)} for i in range(40)]
    funcs = make_funcs(specs)
    vals = orc_eval(funcs, pts)
    ctx.set_mesh(pts, tets)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    compare_ia(ctx, ctx.download_mesh(), orc_run("ia", pts, tets, vals), cnt)


def test_empty_and_tiny_inputs(ctx):
    pts, tets = orc_grid(4)
    far = make_funcs([{"type": "sphere", "center": [9, 9, 9], "radius": 0.5}])
    ctx.set_mesh(pts, tets)
    ctx.set_functions(far)
    cnt = ctx.run()
    assert cnt.num_intersecting_tet == 0 and cnt.num_verts == 0 and cnt.num_faces == 0
    mesh = ctx.download_mesh()
    assert mesh["face_offsets"].tolist() == [0]
    # a single tetrahedron cut by one plane
    p1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float64)
    t1 = np.array([[0, 1, 2, 3]], np.uint32)
    f1 = make_funcs([{"type": "plane", "point": [0.25, 0, 0], "normal": [1, 0, 0]}])
    ctx.set_mesh(p1, t1)
    ctx.set_functions(f1)
    cnt = ctx.run()
    compare_ia(ctx, ctx.download_mesh(), orc_run("ia", p1, t1, orc_eval(f1, p1)), cnt)


def test_negate_flag_matches_csg_negation(ctx, rin):
    """csg(): funcVals * -1 when !positive_inside (src/csg.cpp:37)."""
    pts, tets = orc_grid(12)
    funcs = make_funcs(synthetic_functions("C2"))
    vals = orc_eval(funcs, pts)
    port = orc_run("ia", pts, tets, vals, flags=FLAG_LOOKUP | FLAG_SECONDARY | 4)
    ctx.set_mesh(pts, tets)
    ctx.set_values(vals)
    cnt = ctx.run(flags=FLAG_LOOKUP | FLAG_SECONDARY | rin.FLAG_NEGATE)
    compare_ia(ctx, ctx.download_mesh(), port, cnt)


def test_tet_range_shard_equals_oracle_on_the_same_range(ctx):
    """Contiguous tet ranges (slab sharding): each range reproduces the oracle run on that range."""
    R = 16
    pts, tets = orc_grid(R)
    funcs = make_funcs(synthetic_functions("C2"))
    vals = orc_eval(funcs, pts)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    T = len(tets)
    for first, count in ((0, T // 2), (T // 2, T - T // 2)):
        ctx.set_tet_range(first, count)
        cnt = ctx.run()
        port = orc_run("ia", pts, tets, vals, tet_first=first, tet_count=count)
        compare_ia(ctx, ctx.download_mesh(), port, cnt)


def test_full_size_c2_128(ctx):
    """BASELINE C2 (128^3, 10.5 M tets, 8 functions) at full size: bit-exact against the oracle and
    first-occurrence / uniqueness properties."""
    R = 128
    funcs = make_funcs(synthetic_functions("C2"))
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    mesh = ctx.download_mesh()
    # properties that do not need the oracle
    assert np.all(np.diff(mesh["vert_tet"].astype(np.int64)) >= 0)  # first-occurrence order
    keys = np.concatenate([mesh["vert_simplex_size"][:, None].astype(np.uint32), mesh["vert_simplex"],
                           mesh["vert_funcs"][:, :3]], axis=1)
    shared = keys[mesh["vert_simplex_size"] < 4]
    assert len(np.unique(shared, axis=0)) == len(shared)  # deduplication left no duplicate key
    assert mesh["face_verts"].max() < cnt.num_verts
    assert np.all(np.isfinite(mesh["vert_xyz"])) and np.abs(mesh["vert_xyz"]).max() <= 1.0
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    compare_ia(ctx, mesh, orc_run("ia", pts, tets, vals), cnt)


def test_c4_dense_functions(ctx):
    """BASELINE C4 shape (32 near-coincident spheres, general kernel dominates) at a size the oracle
    finishes in seconds; the 128^3 instance is exercised by the bench."""
    R = 40
    funcs = make_funcs(synthetic_functions("C4"))
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    assert cnt.num_kmore > cnt.num_k1
    compare_ia(ctx, ctx.download_mesh(), orc_run("ia", pts, tets, vals), cnt)


def many_sheet_functions(n=24):
    """n nearly parallel, closely spaced planes: every tet they cross sees all of them (k = n > the 20 functions
    the shared-memory tiers hold) while the complexes stay small (the sheets do not meet inside a tet)."""
    return [{"type": "plane", "point": [0.31 + 0.004 * j, 0.0, 0.0], "normal": [1.0, 0.013 * (j % 5), 0.007 * (j % 3)]}
            for j in range(n)]


def test_more_functions_than_the_shared_memory_tiers_hold(ctx):
    """k = 24 active functions per tet: small tier (<= 4) and mid tier (<= 20) pass, the per-thread big tier runs."""
    R = 4
    funcs = make_funcs(many_sheet_functions(24))
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    port = orc_run("ia", pts, tets, vals)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    cnt = ctx.run()
    assert cnt.num_kmore > 0 and cnt.num_active_funcs >= 24 * cnt.num_kmore // 2
    compare_ia(ctx, ctx.download_mesh(), port, cnt)
