"""One rank of the multi-GPU parity test (launched by tests/test_gpu_nccl.py through torch.distributed.run):
every rank runs its x-slab through the C-ABI, the slab-boundary exchange runs on the device over NCCL
(rin_exchange_nccl), every rank writes its slice of the merged mesh into a POSIX shared-memory segment and
rank 0 compares the merged mesh with the CPU oracle's single-process result, bit for bit."""
import os
import sys
from multiprocessing import shared_memory

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust-implicit-surface-networks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import rin_b200 as rin
    import sharding
    from helpers import make_funcs, orc_eval, orc_grid, orc_run, synthetic_functions

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    ctx = rin.Context(local)
    uid = [rin.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    ctx.nccl_init(uid[0], rank, world)

    # degenerate inputs: iso-faces / material interfaces ON the slab planes (ghost-layer negotiation inside
    # rin_exchange_nccl, src/extract_mesh.cpp:240-253, 833-981)
    special = {"ia_x0": [{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                         {"type": "sphere", "center": [0, 0, 0], "radius": 0.5, "squared": True},
                         {"type": "plane", "point": [0, 0, 0], "normal": [0, 1, 0]}],
               "mi_x0": [{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                         {"type": "plane", "point": [0, 0, 0], "normal": [-1, 0, 0]},
                         {"type": "sphere", "center": [0, 0.5, 0], "radius": 0.6}]}
    cases = [("ia", "C2", 20), ("mi", "C3", 18), ("ia", "C2", 33), ("ia", "ia_x0", 16), ("mi", "mi_x0", 16)]
    for mode, fset, R in cases:
        for allgather in (False, True):
            if allgather:
                os.environ["RIN_X_ALLGATHER"] = "1"  # the general (non-slab) protocol on the same input
            else:
                os.environ.pop("RIN_X_ALLGATHER", None)
            funcs = make_funcs(special[fset] if fset in special else synthetic_functions(fset))
            ctx.generate_grid(R)  # resets the cached vertex windows
            ctx.set_functions(funcs)
            ctx.set_tet_range(*sharding.slab_range(R, rank, world))
            for rep in range(4):  # the second pass reuses the negotiated capacities; passes 3 and 4 go through
                # rin_run_exchange (one synchronisation; fused for IA on slabs, two calls otherwise)
                if rep < 2:
                    cnt = ctx.run(rin.MODE_IA if mode == "ia" else rin.MODE_MI)
                    info = ctx.exchange_nccl()
                else:
                    info = ctx.run_exchange(rin.MODE_IA if mode == "ia" else rin.MODE_MI)
                cnt = ctx.counts()
            layout = sharding.merged_layout(info)
            sizes = {k: int(np.prod(sh)) * np.dtype(dt).itemsize for k, (sh, dt) in layout.items()}
            name = [None]
            if rank == 0:
                shm = shared_memory.SharedMemory(create=True, size=max(1, sum(sizes.values())))
                name[0] = shm.name
            dist.broadcast_object_list(name, 0)
            if rank != 0:
                shm = shared_memory.SharedMemory(name=name[0])
            merged, off = {}, 0
            for k, (sh, dt) in layout.items():
                merged[k] = np.ndarray(sh, dt, buffer=shm.buf, offset=off)
                off += sizes[k]
            views = sharding.slice_views(merged, info, cnt)
            # ranks write overlapping last/first offset entries: lower ranks first
            for r in range(world):
                if r == rank:
                    ctx.download_mesh(views)  # offsets arrive rebased into the merged arrays
                dist.barrier()
            if rank == 0:
                pts, tets = orc_grid(R)
                ref = orc_run(mode, pts, tets, orc_eval(funcs, pts))
                assert info["n_verts_total"] == ref["stats"][7] and info["n_faces_total"] == ref["stats"][8]
                assert np.array_equal(merged["vert_xyz"], ref["vert_xyz"].reshape(-1, 3))
                assert np.array_equal(merged["face_offsets"].astype(np.int64), ref["face_offsets"])
                assert np.array_equal(merged["face_verts"].astype(np.int64), ref["face_verts"])
                assert np.array_equal(merged["face_tet_offsets"].astype(np.int64), ref["face_tet_offsets"])
                assert np.array_equal(merged["face_tets"].astype(np.int64).ravel(), ref["face_tets"])
                ff = merged["face_funcs"].astype(np.int64)
                ff[ff == 0xFFFFFFFF] = -1
                assert np.array_equal(ff.ravel(), ref["face_funcs"])
                nrec = 11 if mode == "mi" else 10
                rec = ref["vert_rec"].reshape(-1, nrec)
                assert np.array_equal(merged["vert_tet"].astype(np.int64), rec[:, 0])
                assert np.array_equal(merged["vert_simplex_size"].astype(np.int64), rec[:, 2])
                print("nccl parity ok:", mode, fset, R, "allgather" if allgather else "neighbours",
                      info["n_verts_total"], info["n_faces_total"], flush=True)
            dist.barrier()
            del merged, views
            shm.close()
            if rank == 0:
                shm.unlink()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
