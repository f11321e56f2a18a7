"""CPU (gloo, world_size 2): the slab-boundary exchange protocol of sharding.py with the oracle as
the per-rank engine; the merged result must equal the single-process oracle run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust-implicit-surface-networks_b200"))


def ia_keys_from_vert_rec(rec):
    """(sorted simplex verts, f0 | f1 << 16) as on the device (cand_key of kernels_ia.cuh)."""
    rec = rec.reshape(-1, 10)
    sv = rec[:, 3:6].copy()
    sv[sv < 0] = 0xFFFFFFFF
    f0 = np.where(rec[:, 7] < 0, 0xFFFF, rec[:, 7])
    f1 = np.where(rec[:, 8] < 0, 0xFFFF, rec[:, 8])
    return np.concatenate([sv, (f0 | (f1 << 16))[:, None]], axis=1).astype(np.uint32), rec[:, 2]


def worker(rank, world, port, R, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_funcs, orc_eval, orc_grid, orc_run, synthetic_functions
    import sharding
    pts, tets = orc_grid(R)
    vals = orc_eval(make_funcs(synthetic_functions("C2")), pts)
    first, count = sharding.slab_range(R, rank, world)
    b = orc_run("ia", pts, tets, vals, tet_first=first, tet_count=count)
    keys, sizes = ia_keys_from_vert_rec(b["vert_rec"])
    used = tets[first:first + count]
    eng = sharding.NumpyEngine(keys, sizes, int(used.min()), int(used.max()))
    info = sharding.exchange(eng, rank, world, sharding.torch_gather(dist), len(b["face_offsets"]) - 1)
    own = eng.own_idx >= 0
    np.savez(out % rank, xyz=b["vert_xyz"].reshape(-1, 3)[own], face_verts=eng.gid[b["face_verts"]],
             face_offsets=b["face_offsets"], vert_offset=info["vert_offset"], n_total=info["n_verts_total"])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_exchange_reproduces_single_run(tmp_path, world):
    R = 12
    out = str(tmp_path / "rank%d.npz")
    port = 29500 + (os.getpid() % 1000) + world
    mp.spawn(worker, args=(world, port, R, out), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_funcs, orc_eval, orc_grid, orc_run, synthetic_functions
    pts, tets = orc_grid(R)
    vals = orc_eval(make_funcs(synthetic_functions("C2")), pts)
    ref = orc_run("ia", pts, tets, vals)
    parts = [np.load(out % r) for r in range(world)]
    xyz = np.concatenate([p["xyz"] for p in parts])
    fv = np.concatenate([p["face_verts"] for p in parts])
    assert int(parts[0]["n_total"]) == len(xyz) == ref["stats"][7]
    assert np.array_equal(xyz, ref["vert_xyz"].reshape(-1, 3))  # same vertices, same global order
    assert np.array_equal(fv, ref["face_verts"])                # faces refer to the same global ids
    sizes = np.concatenate([np.diff(p["face_offsets"]) for p in parts])
    assert np.array_equal(sizes, np.diff(ref["face_offsets"]))


def test_slabs_of_equal_cost_and_balanced_planes():
    """Slab plans: every rank gets at least one cube plane, boundaries ascend, the largest slab cost is minimal
    (checked against brute force on a small case)."""
    import itertools
    import sharding
    rng = np.random.default_rng(5)
    cost = rng.random(12) + 0.05
    for world in (1, 2, 3, 5):
        plan = sharding.slabs_of_equal_cost(cost, world)
        assert plan[0] == 0 and plan[-1] == 12 and all(b > a for a, b in zip(plan, plan[1:])) and len(plan) == world + 1
        worst = max(cost[a:b].sum() for a, b in zip(plan, plan[1:]))
        best = min(max(cost[a:b].sum() for a, b in zip((0,) + cuts, cuts + (12,)))
                   for cuts in itertools.combinations(range(1, 12), world - 1))
        assert worst <= best * (1 + 1e-9)
    assert sharding.slabs_of_equal_cost(np.ones(3), 5)[-1] == 3  # more ranks than planes: trailing ranks are empty
    hist = np.zeros(16)
    hist[6:10] = 1000.0  # the surface sits in the middle: the middle slabs get fewer planes
    plan = sharding.balanced_slab_planes(hist, 4)
    widths = np.diff(plan)
    assert plan[0] == 0 and plan[-1] == 16 and widths.min() >= 1 and widths[0] > widths[1]
