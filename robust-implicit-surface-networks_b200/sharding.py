"""Slab sharding across GPUs: contiguous tet ranges per rank and the slab-boundary key exchange.

Tets are independent through evaluation, filter, per-tet arrangement and candidate emission; the
only cross-rank step is the identification of vertices found by two ranks on their shared vertex
plane.  The rank holding the LOWER tet range owns such a vertex (it is the first occurrence in
tet order, the numbering rule of /root/reference/src/extract_mesh.cpp:139,169,214), so that the
concatenation of the ranks' vertex arrays equals the single-process result.

`exchange()` is engine-agnostic: `engine` is a rin_b200.Context (device kernels behind the C-ABI)
or any object with the same four methods; `gather(array) -> [array per rank]` is an all-gather
(torch.distributed NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def slab_range(R, rank, world):
    """Tet range of x-slabs [i0, i1) of a generated grid (cubes are i-major, src/io.cpp:122-125)."""
    i0, i1 = R * rank // world, R * (rank + 1) // world
    per_slab = 5 * R * R
    return i0 * per_slab, (i1 - i0) * per_slab


def slab_of_planes(R, i0, i1):
    """Tet range of the x-slabs [i0, i1)."""
    per_slab = 5 * R * R
    return i0 * per_slab, (i1 - i0) * per_slab


# cost model of one pass (seconds): streaming stages per tet + active-tet stages per active tet
# (B200, BASELINE C2: evaluation + filter 0.09 ms for 10.5 M tets, the rest 0.3 ms for 0.68 M active tets)
COST_PER_TET, COST_PER_ACTIVE = 8.6e-12, 4.5e-10


def balanced_slab_planes(active_per_plane, world):
    """Boundaries 0 = b_0 < b_1 < ... < b_world = R of x-slabs that minimise the largest estimated cost, given the
    number of active tets in every x-plane of cubes (from a calibration pass).  The surface is rarely spread
    evenly along x, and every rank waits for the slowest one in the exchange."""
    a = np.asarray(active_per_plane, np.float64)
    R = len(a)
    if world >= R:
        return [min(i, R) for i in range(world + 1)][:world] + [R]
    cost = COST_PER_TET * 5 * R * R + COST_PER_ACTIVE * a
    pre = np.concatenate([[0.0], np.cumsum(cost)])

    def plan(limit):
        b, i = [0], 0
        for s in range(world):
            left = world - s - 1  # slabs still to come need one plane each
            j = i + 1
            while j < R - left and pre[j + 1] - pre[i] <= limit:
                j += 1
            b.append(j)
            i = j
        return b if b[-1] == R else None

    lo, hi = float(cost.max()), float(pre[-1])
    best = plan(hi)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        p = plan(mid)
        if p is not None:
            best, hi = p, mid
        else:
            lo = mid
    return best


def slabs_of_equal_cost(cost_per_plane, world):
    """Boundaries 0 = b_0 < ... < b_world = R that minimise the largest slab cost for a given cost of every cube
    plane (bisection on the limit, greedy filling; every slab keeps at least one plane)."""
    cost = np.asarray(cost_per_plane, np.float64)
    R = len(cost)
    if world >= R:
        return [min(i, R) for i in range(world + 1)][:world] + [R]
    pre = np.concatenate([[0.0], np.cumsum(cost)])

    def plan(limit):
        b, i = [0], 0
        for s in range(world):
            left = world - s - 1
            j = i + 1
            while j < R - left and pre[j + 1] - pre[i] <= limit:
                j += 1
            b.append(j)
            i = j
        return b if b[-1] == R else None

    lo, hi = float(cost.max()), float(pre[-1])
    best = plan(hi)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        p = plan(mid)
        if p is not None:
            best, hi = p, mid
        else:
            lo = mid
    return best


class ExchangeState:
    """Per-run cache: the shared vertex window does not change between passes over the same mesh."""

    def __init__(self):
        self.window = None


def exchange(engine, rank, world, gather, n_faces, state=None, degenerate=0, ghost_rerun=None):
    """Runs the boundary protocol (two all-gathers per pass).  Returns dict(vert_offset,
    face_offset, n_verts_total, n_faces_total, n_own).

    Degenerate inputs (a function vanishing at grid vertices, materials tying there) can put an iso-face on a
    tet face of the slab plane, where the reference pairs the two incident tets (src/extract_mesh.cpp:240-253) or
    matches their materials (:833-981).  `degenerate` is this rank's num_degenerate_vertex; when any rank reports
    one, every rank calls `ghost_rerun()` (rin_set_ghost_tets with one cube layer + rin_run; returns the new face
    count) before the exchange - the same negotiation rin_exchange_nccl does on the device."""
    if state is None:
        state = ExchangeState()
    if ghost_rerun is not None:
        flags = gather(np.array([1 if degenerate else 0], np.int64))
        if any(int(f[0]) for f in flags):
            n_faces = ghost_rerun()
            state.window = None
    if state.window is None:
        lo, hi = engine.vertex_range()
        ranges = gather(np.array([lo, hi], np.int64))
        # window covering every overlap of this rank's vertex range with another rank's
        w_lo, w_hi = None, None
        for s, (slo, shi) in enumerate(ranges):
            if s == rank:
                continue
            a, b = max(lo, int(slo)), min(hi, int(shi))
            if a <= b:
                w_lo = a if w_lo is None else min(w_lo, a)
                w_hi = b if w_hi is None else max(w_hi, b)
        state.window = (1, 0) if w_lo is None else (w_lo, w_hi)
    w_lo, w_hi = state.window
    # 1. keys of every candidate in the shared window; a key found by a lower rank is foreign
    keys, _ = engine.boundary_export(False, w_lo, w_hi)
    all_keys = gather(keys.reshape(-1).astype(np.int64))
    lower = [k.reshape(-1, 4) for k in all_keys[:rank]]
    lower = np.concatenate(lower).astype(np.uint32) if lower else np.zeros((0, 4), np.uint32)
    n_own = engine.mark_foreign(lower)
    # 2. (n_own, n_faces) + owned shared keys with their own index; receivers add the owner's offset
    keys2, own_idx = engine.boundary_export(True, w_lo, w_hi)
    msg = np.concatenate([np.array([n_own, n_faces], np.int64),
                          np.concatenate([keys2.astype(np.int64), own_idx.astype(np.int64)[:, None]],
                                         axis=1).reshape(-1)])
    allm = gather(msg)
    n_owns = [int(m[0]) for m in allm]
    n_fs = [int(m[1]) for m in allm]
    offsets = np.concatenate([[0], np.cumsum(n_owns)])
    low = []
    for s in range(rank):
        p = allm[s][2:].reshape(-1, 5).copy()
        p[:, 4] += offsets[s]
        low.append(p)
    low = np.concatenate(low) if low else np.zeros((0, 5), np.int64)
    vert_offset = int(offsets[rank])
    engine.finalize_sharded(vert_offset, low[:, :4].astype(np.uint32), low[:, 4].astype(np.uint32))
    return {"vert_offset": vert_offset, "face_offset": int(sum(n_fs[:rank])), "n_own": n_own,
            "n_verts_total": int(offsets[-1]), "n_faces_total": int(sum(n_fs))}


def torch_gather(dist, device=None):
    """all-gather of variable-length int64 vectors over torch.distributed.  Buffers are padded to a
    remembered capacity so that the usual pass needs ONE collective; the element count travels in
    slot 0 and a pass whose payload outgrew the capacity is repeated with a larger one."""
    import torch

    cap = {"n": 1024}

    def gather(arr):
        arr = np.ascontiguousarray(arr, np.int64).reshape(-1)
        world = dist.get_world_size()
        while True:
            m = cap["n"]
            host = np.zeros(m + 1, np.int64)
            host[0] = arr.size
            if arr.size <= m:
                host[1:1 + arr.size] = arr
            buf = torch.from_numpy(host).to(device) if device is not None else torch.from_numpy(host)
            out = torch.empty((world, m + 1), dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(out, buf) if hasattr(dist, "all_gather_into_tensor") and device is not None \
                else dist.all_gather(list(out.unbind(0)), buf)
            res = out.cpu().numpy()
            need = int(res[:, 0].max())
            if need <= m:
                return [res[r, 1:1 + int(res[r, 0])] for r in range(world)]
            cap["n"] = int(need * 1.25) + 1024

    return gather


class NumpyEngine:
    """The same four steps on host arrays (used by the CPU tests with the oracle as per-rank engine)."""

    def __init__(self, keys, sizes, v_lo, v_hi):
        self.keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 4)  # one per local vertex
        self.sizes = np.asarray(sizes)
        self.lo, self.hi = v_lo, v_hi
        self.own_idx = None
        self.gid = None

    def vertex_range(self):
        return self.lo, self.hi

    def _in_window(self, lo, hi):
        k, s = self.keys.astype(np.int64), self.sizes
        ok = (s < 4) & (k[:, 0] >= lo) & (k[:, 0] <= hi)
        ok &= (s < 2) | ((k[:, 1] >= lo) & (k[:, 1] <= hi))
        ok &= (s < 3) | ((k[:, 2] >= lo) & (k[:, 2] <= hi))
        return ok

    def boundary_export(self, own_only, lo, hi):
        sel = self._in_window(lo, hi)
        if own_only:
            sel &= self.own_idx >= 0
            return self.keys[sel], self.own_idx[sel].astype(np.uint32)
        return self.keys[sel], np.nonzero(sel)[0].astype(np.uint32)

    def mark_foreign(self, keys):
        foreign = {tuple(k) for k in np.asarray(keys).reshape(-1, 4).tolist()}
        own = np.array([not (s < 4 and tuple(k) in foreign) for k, s in zip(self.keys.tolist(), self.sizes)],
                       bool)
        self.own_idx = np.where(own, np.cumsum(own) - 1, -1)
        return int(own.sum())

    def finalize_sharded(self, offset, keys, gids):
        table = {tuple(k): int(g) for k, g in zip(np.asarray(keys).reshape(-1, 4).tolist(), gids)}
        self.gid = np.array([offset + o if o >= 0 else table[tuple(k)]
                             for o, k in zip(self.own_idx, self.keys.tolist())], np.int64)


MESH_VERT_KEYS = ("vert_tet", "vert_local", "vert_simplex_size", "vert_simplex", "vert_funcs", "vert_xyz")


def merged_layout(info):
    """Sizes of the merged mesh arrays from one rank's exchange result (identical on every rank)."""
    nv, nf, nfv, nft = info["n_verts_total"], info["n_faces_total"], info["n_fv_total"], info["n_ft_total"]
    return {"vert_tet": ((nv,), np.uint32), "vert_local": ((nv,), np.uint8), "vert_simplex_size": ((nv,), np.uint8),
            "vert_simplex": ((nv, 4), np.uint32), "vert_funcs": ((nv, 4), np.uint32), "vert_xyz": ((nv, 3), np.float64),
            "face_offsets": ((nf + 1,), np.uint32), "face_verts": ((nfv,), np.uint32),
            "face_tet_offsets": ((nf + 1,), np.uint32), "face_tets": ((nft, 2), np.uint32),
            "face_funcs": ((nf, 2), np.uint32)}


def slice_views(merged, info, counts):
    """Views of this rank's slice inside the merged arrays (what rin_download_mesh writes into).  After
    rin_exchange_nccl the offset arrays arrive rebased into the merged arrays."""
    v0, f0, fv0, ft0 = info["vert_offset"], info["face_offset"], info["fv_offset"], info["ft_offset"]
    nv, nf, nfv, nft = counts.num_verts, counts.num_faces, counts.num_face_verts, counts.num_face_tets
    out = {k: merged[k][v0:v0 + nv] for k in MESH_VERT_KEYS}
    out["face_verts"] = merged["face_verts"][fv0:fv0 + nfv]
    out["face_tets"] = merged["face_tets"][ft0:ft0 + nft]
    out["face_funcs"] = merged["face_funcs"][f0:f0 + nf]
    # every slice writes nf + 1 offsets; the last one of a slice is overwritten by its successor's first
    out["face_offsets"] = merged["face_offsets"][f0:f0 + nf + 1]
    out["face_tet_offsets"] = merged["face_tet_offsets"][f0:f0 + nf + 1]
    return out
