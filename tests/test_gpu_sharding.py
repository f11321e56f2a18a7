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
