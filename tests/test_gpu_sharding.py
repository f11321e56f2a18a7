"""GPU: slab sharding with the device-side boundary exchange (rin_boundary_export /
rin_mark_foreign / rin_finalize_sharded).  Ranks are simulated by contexts on one GPU driven from
threads; the all-gather is an in-process rendezvous.  The merged mesh must equal the oracle's
single-process result bit for bit."""
import threading

import numpy as np
import pytest

from helpers import make_funcs, orc_eval, orc_grid, orc_run, synthetic_functions

pytestmark = pytest.mark.gpu


class LocalGather:
    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.barrier = threading.Barrier(world)

    def for_rank(self, rank):
        def gather(arr):
            self.slots[rank] = np.array(arr, np.int64).reshape(-1).copy()
            self.barrier.wait()
            out = [s.copy() for s in self.slots]
            self.barrier.wait()
            return out
        return gather


@pytest.mark.parametrize("world,R,mode", [(2, 16, "ia"), (4, 24, "ia"), (3, 18, "mi")])
def test_sharded_run_equals_single_run(world, R, mode):
    import rin_b200 as rin
    import sharding
    cfg = "C2" if mode == "ia" else "C3"
    funcs = make_funcs(synthetic_functions(cfg))
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    ref = orc_run(mode, pts, tets, vals)
    lg = LocalGather(world)
    results = [None] * world
    errors = []

    def rank_main(rank):
        try:
            ctx = rin.Context(0)
            ctx.generate_grid(R)
            ctx.set_functions(funcs)
            first, count = sharding.slab_range(R, rank, world)
            ctx.set_tet_range(first, count)
            cnt = ctx.run(rin.MODE_IA if mode == "ia" else rin.MODE_MI)
            info = sharding.exchange(ctx, rank, world, lg.for_rank(rank), cnt.num_faces)
            mesh = ctx.download_mesh()
            results[rank] = (info, mesh)
            ctx.close()
        except Exception as ex:  # surface the failure in the main thread
            errors.append(ex)
            lg.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    xyz = np.concatenate([m["vert_xyz"] for _, m in results])
    fv = np.concatenate([m["face_verts"] for _, m in results]).astype(np.int64)
    ft = np.concatenate([m["face_tets"] for _, m in results]).astype(np.int64)
    assert results[0][0]["n_verts_total"] == len(xyz) == ref["stats"][7]
    assert np.array_equal(xyz, ref["vert_xyz"].reshape(-1, 3))
    assert np.array_equal(fv, ref["face_verts"])
    assert np.array_equal(ft.ravel(), ref["face_tets"])
    assert sum(m["face_offsets"].shape[0] - 1 for _, m in results) == ref["stats"][8]


# Degenerate inputs whose iso-faces / material interfaces lie on tet faces of the slab planes (even R, cuts at grid
# planes through the origin): the reference pairs the two tets of such a face (src/extract_mesh.cpp:240-253) and
# matches materials across it (:833-981).  Sharded runs negotiate one ghost cube layer (rin_set_ghost_tets).
DEGENERATE = {
    "ia_plane_x0": ("ia", [{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                           {"type": "sphere", "center": [0, 0, 0], "radius": 0.5, "squared": True},
                           {"type": "plane", "point": [0, 0, 0], "normal": [0, 1, 0]}]),
    "ia_planes_at_cuts": ("ia", [{"type": "plane", "point": [-0.5, 0, 0], "normal": [1, 0, 0]},
                                 {"type": "plane", "point": [0, 0, 0], "normal": [-1, 0, 0]},
                                 {"type": "plane", "point": [0.5, 0, 0], "normal": [1, 0, 0]},
                                 {"type": "sphere", "center": [0.1, 0, 0], "radius": 0.7, "squared": False}]),
    "mi_x_vs_negx": ("mi", [{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                            {"type": "plane", "point": [0, 0, 0], "normal": [-1, 0, 0]}]),
    "mi_sym_spheres": ("mi", [{"type": "sphere", "center": [-0.5, 0, 0], "radius": 0.7},
                              {"type": "sphere", "center": [0.5, 0, 0], "radius": 0.7},
                              {"type": "sphere", "center": [0, 0.5, 0], "radius": 0.6},
                              {"type": "sphere", "center": [0, -0.5, 0], "radius": 0.6}]),
}


@pytest.mark.parametrize("name", sorted(DEGENERATE))
@pytest.mark.parametrize("world,R", [(2, 8), (4, 16)])
def test_sharded_degenerate_faces_on_the_slab_plane(name, world, R):
    import rin_b200 as rin
    import sharding
    mode, specs = DEGENERATE[name]
    funcs = make_funcs(specs)
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    ref = orc_run(mode, pts, tets, vals)
    assert ref.error == ""
    lg = LocalGather(world)
    results = [None] * world
    errors = []

    def rank_main(rank):
        try:
            ctx = rin.Context(0)
            ctx.generate_grid(R)
            ctx.set_functions(funcs)
            ctx.set_tet_range(*sharding.slab_range(R, rank, world))
            m = rin.MODE_IA if mode == "ia" else rin.MODE_MI
            cnt = ctx.run(m)

            def ghost_rerun():
                ctx.set_ghost_tets(5 * R * R, 5 * R * R)
                return ctx.run(m).num_faces

            info = sharding.exchange(ctx, rank, world, lg.for_rank(rank), cnt.num_faces,
                                     degenerate=cnt.num_degenerate_vertex, ghost_rerun=ghost_rerun)
            results[rank] = (info, ctx.download_mesh())
            ctx.close()
        except Exception as ex:
            errors.append(ex)
            lg.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    cat = lambda k: np.concatenate([m[k] for _, m in results])
    assert np.array_equal(cat("vert_xyz"), ref["vert_xyz"].reshape(-1, 3))
    assert np.array_equal(cat("face_verts").astype(np.int64), ref["face_verts"])
    assert np.array_equal(cat("face_tets").astype(np.int64).ravel(), ref["face_tets"])
    ff = cat("face_funcs").astype(np.int64)
    ff[ff == 0xFFFFFFFF] = -1
    assert np.array_equal(ff.ravel(), ref["face_funcs"])
    # per-face sizes (offsets are rank-local here: compare the differences)
    sizes = np.concatenate([np.diff(m["face_offsets"].astype(np.int64)) for _, m in results])
    assert np.array_equal(sizes, np.diff(ref["face_offsets"]))
    tsz = np.concatenate([np.diff(m["face_tet_offsets"].astype(np.int64)) for _, m in results])
    assert np.array_equal(tsz, np.diff(ref["face_tet_offsets"]))
    assert tsz.max() == 2 or mode == "mi"  # IA: some face is shared by two tets
    nrec = 11 if mode == "mi" else 10
    rec = ref["vert_rec"].reshape(-1, nrec)
    assert np.array_equal(cat("vert_tet").astype(np.int64), rec[:, 0])
    assert np.array_equal(cat("vert_simplex_size").astype(np.int64), rec[:, 2])
