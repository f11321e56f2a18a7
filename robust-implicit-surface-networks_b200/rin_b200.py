"""ctypes binding of librin_b200.so (the C-ABI declared in include/rin_b200.h).

Used by tests/, bench.py and __graft_entry__.py.  There is no CPU path behind this module: if the
CUDA library is missing it raises, and every computing call needs a CUDA device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librin_b200.so")

MODE_IA, MODE_MI = 0, 1
FLAG_LOOKUP, FLAG_SECONDARY, FLAG_NEGATE = 1, 2, 4

FUNC_DESC = np.dtype([("type", "<i4"), ("flip", "<i4"), ("p", "<f8", (10,))])


class Counts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "num_pts", "num_tets", "num_funcs", "num_degenerate_vertex", "num_intersecting_tet",
        "num_k1", "num_k2", "num_kmore", "num_verts", "num_faces", "num_face_verts",
        "num_face_tets", "num_general_tets", "num_exact_fallbacks", "num_active_funcs")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class MeshOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "vert_tet", "vert_local", "vert_simplex_size", "vert_simplex", "vert_funcs", "vert_xyz",
        "face_offsets", "face_verts", "face_tet_offsets", "face_tets", "face_funcs")]


class RinError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rin_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "librin_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.rin_last_error.restype = C.c_char_p
        L.rin_stage_name.restype = C.c_char_p
        L.rin_stage_name.argtypes = [C.c_int]
        L.rin_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.rin_destroy.argtypes = [C.c_void_p]
        L.rin_set_mesh_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int]
        L.rin_generate_grid.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.rin_set_mesh_host_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                                              C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
        L.rin_set_values_host_range.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
        L.rin_set_tet_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.rin_set_ghost_tets.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.rin_set_functions.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.rin_set_values_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        L.rin_run.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.rin_get_counts.argtypes = [C.c_void_p, C.POINTER(Counts)]
        L.rin_tet_maps.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.rin_download_tet_maps.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.rin_download_mesh.argtypes = [C.c_void_p, C.POINTER(MeshOut)]
        L.rin_download_active.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rin_download_values.argtypes = [C.c_void_p, C.c_void_p]
        L.rin_download_active_tets.argtypes = [C.c_void_p, C.c_void_p]
        L.rin_download_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rin_get_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.rin_set_stage_timing.argtypes = [C.c_void_p, C.c_int]
        L.rin_get_launch_count.argtypes = [C.c_void_p]
        L.rin_get_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.POINTER(C.c_float)]
        L.rin_boundary_export.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_uint64, C.POINTER(C.c_uint64)]
        L.rin_mark_foreign.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.rin_finalize_sharded.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        L.rin_get_vertex_range.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.rin_nccl_unique_id.argtypes = [C.c_void_p]
        L.rin_nccl_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.rin_exchange_nccl.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4
        L.rin_get_exchange_time.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.rin_get_exchange_parts.argtypes = [C.c_void_p, C.c_void_p]
        L.rin_run_exchange.argtypes = [C.c_void_p, C.c_int, C.c_uint32] + [C.POINTER(C.c_uint64)] * 4
        L.rin_get_exchange_offsets.argtypes = [C.c_void_p, C.c_void_p]
        L.rin_robust_test.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.rin_run_host.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p,
                                   C.c_uint64, C.c_int, C.c_void_p, C.c_uint32, C.POINTER(Counts)]
        L.rin_get_complexes.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p,
                                        C.c_void_p, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def nccl_unique_id():
    uid = np.zeros(128, np.uint8)
    rc = lib().rin_nccl_unique_id(uid.ctypes.data)
    if rc != 0:
        raise RinError(rc, lib().rin_last_error().decode())
    return uid


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Context:
    """One engine instance bound to one GPU (mirrors rin_ctx)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._check(lib().rin_create(device, C.byref(self._h)))
        self._keep = []

    def _check(self, rc):
        if rc != 0:
            raise RinError(rc, lib().rin_last_error().decode())

    def close(self):
        if self._h:
            lib().rin_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs
    def set_mesh(self, pts, tets):
        pts = np.ascontiguousarray(pts, np.float64)
        if tets.dtype not in (np.uint32, np.uint64):
            tets = tets.astype(np.uint64)
        tets = np.ascontiguousarray(tets)
        self._check(lib().rin_set_mesh_host(self._h, pts.ctypes.data, len(pts), tets.ctypes.data, len(tets),
                                            tets.dtype.itemsize))

    def set_mesh_range(self, n_pts, n_tets, pts_slice, v_first, tets_slice, t_first):
        assert tets_slice.dtype in (np.uint32, np.uint64) and pts_slice.dtype == np.float64
        self._check(lib().rin_set_mesh_host_range(self._h, n_pts, n_tets, pts_slice.ctypes.data, v_first,
                                                  len(pts_slice), tets_slice.ctypes.data, t_first, len(tets_slice),
                                                  tets_slice.dtype.itemsize))

    def set_values_range(self, vals_slice, v_first):
        assert vals_slice.dtype == np.float64 and vals_slice.flags.c_contiguous
        self._check(lib().rin_set_values_host_range(self._h, vals_slice.ctypes.data, v_first, vals_slice.shape[0],
                                                    vals_slice.shape[1]))

    def generate_grid(self, R, bmin=(-1, -1, -1), bmax=(1, 1, 1)):
        a = np.asarray(bmin, np.float64)
        b = np.asarray(bmax, np.float64)
        self._check(lib().rin_generate_grid(self._h, R, a.ctypes.data, b.ctypes.data))
        self.grid_R = R

    def set_tet_range(self, first, count):
        self._check(lib().rin_set_tet_range(self._h, first, count))

    def set_ghost_tets(self, below, above):
        """Ghost tets of a sharded run on degenerate inputs (include/rin_b200.h); call after set_tet_range."""
        self._check(lib().rin_set_ghost_tets(self._h, below, above))

    def set_functions(self, funcs):
        funcs = np.ascontiguousarray(funcs, FUNC_DESC)
        self._check(lib().rin_set_functions(self._h, funcs.ctypes.data, len(funcs)))

    def set_values(self, vals):
        vals = np.ascontiguousarray(vals, np.float64)
        self._check(lib().rin_set_values_host(self._h, vals.ctypes.data, vals.shape[0], vals.shape[1]))

    # ---- run
    def run(self, mode=MODE_IA, flags=FLAG_LOOKUP | FLAG_SECONDARY):
        self._check(lib().rin_run(self._h, mode, flags))
        return self.counts()

    def run_host(self, mode, flags, pts, tets, vals):
        cnt = Counts()
        self._check(lib().rin_run_host(self._h, mode, flags, pts.ctypes.data, len(pts), tets.ctypes.data,
                                       len(tets), tets.dtype.itemsize, vals.ctypes.data, vals.shape[1],
                                       C.byref(cnt)))
        return cnt

    def counts(self):
        cnt = Counts()
        self._check(lib().rin_get_counts(self._h, C.byref(cnt)))
        return cnt

    def download_mesh(self, into=None):
        n = self.counts()
        if into is None:
            into = {
                "vert_tet": np.empty(n.num_verts, np.uint32),
                "vert_local": np.empty(n.num_verts, np.uint8),
                "vert_simplex_size": np.empty(n.num_verts, np.uint8),
                "vert_simplex": np.empty((n.num_verts, 4), np.uint32),
                "vert_funcs": np.empty((n.num_verts, 4), np.uint32),
                "vert_xyz": np.empty((n.num_verts, 3), np.float64),
                "face_offsets": np.empty(n.num_faces + 1, np.uint32),
                "face_verts": np.empty(n.num_face_verts, np.uint32),
                "face_tet_offsets": np.empty(n.num_faces + 1, np.uint32),
                "face_tets": np.empty((n.num_face_tets, 2), np.uint32),
                "face_funcs": np.empty((n.num_faces, 2), np.uint32),
            }
        mo = MeshOut(**{k: _ptr(v) for k, v in into.items()})
        self._check(lib().rin_download_mesh(self._h, C.byref(mo)))
        return into

    def download_active(self):
        n = self.counts()
        fit = np.empty(n.num_active_funcs, np.uint32)
        start = np.empty(n.num_tets + 1, np.uint64)
        self._check(lib().rin_download_active(self._h, fit.ctypes.data, start.ctypes.data))
        return fit, start

    def download_active_tets(self):
        out = np.empty(self.counts().num_intersecting_tet, np.uint32)
        self._check(lib().rin_download_active_tets(self._h, out.ctypes.data))
        return out

    def download_values(self):
        n = self.counts()
        out = np.empty((n.num_pts, n.num_funcs), np.float64)
        self._check(lib().rin_download_values(self._h, out.ctypes.data))
        return out

    def download_grid(self, V, T):
        pts = np.empty((V, 3), np.float64)
        tets = np.empty((T, 4), np.uint32)
        self._check(lib().rin_download_grid(self._h, pts.ctypes.data, tets.ctypes.data))
        return pts, tets

    # ---- slab-boundary exchange (see sharding.py for the protocol)
    def vertex_range(self):
        lo, hi = C.c_uint32(), C.c_uint32()
        self._check(lib().rin_get_vertex_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def boundary_export(self, own_only, lo, hi):
        n = C.c_uint64()
        self._check(lib().rin_boundary_export(self._h, int(own_only), lo, hi, None, None, 0, C.byref(n)))
        keys = np.empty((n.value, 4), np.uint32)
        ids = np.empty(n.value, np.uint32)
        if n.value:
            self._check(lib().rin_boundary_export(self._h, int(own_only), lo, hi, keys.ctypes.data, ids.ctypes.data,
                                                  n.value, C.byref(n)))
        return keys, ids

    def mark_foreign(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 4)
        n = C.c_uint64()
        self._check(lib().rin_mark_foreign(self._h, keys.ctypes.data if len(keys) else None, len(keys), C.byref(n)))
        return n.value

    def finalize_sharded(self, offset, keys, gids):
        keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 4)
        gids = np.ascontiguousarray(gids, np.uint32)
        self._check(lib().rin_finalize_sharded(self._h, offset, keys.ctypes.data if len(keys) else None,
                                               gids.ctypes.data if len(keys) else None, len(keys)))

    def nccl_init(self, unique_id, rank, world):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert uid.size == 128
        self._check(lib().rin_nccl_init(self._h, uid.ctypes.data, rank, world))

    def exchange_nccl(self):
        """Device-side slab-boundary exchange on the context's stream (neighbour send/recv for slabs, all-gather
        otherwise); returns this rank's offsets in the merged mesh."""
        v = [C.c_uint64() for _ in range(4)]
        self._check(lib().rin_exchange_nccl(self._h, *[C.byref(x) for x in v]))
        o = (C.c_uint64 * 8)()
        self._check(lib().rin_get_exchange_offsets(self._h, o))
        return {"vert_offset": v[0].value, "n_verts_total": v[1].value, "face_offset": v[2].value,
                "n_faces_total": v[3].value, "fv_offset": o[4], "n_fv_total": o[5], "ft_offset": o[6],
                "n_ft_total": o[7]}

    def run_exchange(self, mode=MODE_IA, flags=FLAG_LOOKUP | FLAG_SECONDARY):
        """rin_run + rin_exchange_nccl with one host synchronisation (fused when the library can, see the header)."""
        v = [C.c_uint64() for _ in range(4)]
        self._check(lib().rin_run_exchange(self._h, mode, flags, *[C.byref(x) for x in v]))
        o = (C.c_uint64 * 8)()
        self._check(lib().rin_get_exchange_offsets(self._h, o))
        return {"vert_offset": v[0].value, "n_verts_total": v[1].value, "face_offset": v[2].value,
                "n_faces_total": v[3].value, "fv_offset": o[4], "n_fv_total": o[5], "ft_offset": o[6],
                "n_ft_total": o[7]}

    def exchange_time(self):
        ms = C.c_float()
        self._check(lib().rin_get_exchange_time(self._h, C.byref(ms)))
        return ms.value

    def exchange_parts(self):
        ms = (C.c_float * 10)()
        self._check(lib().rin_get_exchange_parts(self._h, ms))
        return [float(x) for x in ms]

    def get_complexes(self, mode, tet_ids):
        """Full per-tet complexes (layout in include/rin_b200.h) -> (offsets, words)."""
        ids = np.ascontiguousarray(tet_ids, np.uint64)
        n = C.c_uint64()
        off = np.zeros(len(ids) + 1, np.uint64)
        self._check(lib().rin_get_complexes(self._h, mode, 0, ids.ctypes.data, len(ids), off.ctypes.data, None,
                                            C.byref(n)))
        words = np.zeros(max(n.value, 1), np.uint32)
        self._check(lib().rin_get_complexes(self._h, mode, 0, ids.ctypes.data, len(ids), off.ctypes.data,
                                            words.ctypes.data, C.byref(n)))
        return off, words[:n.value]

    def tet_maps(self):
        """Cell-grouping maps of the last IA run (second extract_iso_mesh overload, src/extract_mesh.cpp:268-566)."""
        na, nv, nf = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(lib().rin_tet_maps(self._h, C.byref(na), C.byref(nv), C.byref(nf)))
        out = {"active_tets": np.zeros(na.value, np.uint32), "vert_offsets": np.zeros(na.value + 1, np.uint32),
               "vert_ids": np.zeros(nv.value, np.int64), "face_offsets": np.zeros(na.value + 1, np.uint32),
               "face_ids": np.zeros(nf.value, np.uint32)}
        self._check(lib().rin_download_tet_maps(self._h, out["active_tets"].ctypes.data,
                                                out["vert_offsets"].ctypes.data, out["vert_ids"].ctypes.data,
                                                out["face_offsets"].ctypes.data, out["face_ids"].ctypes.data))
        return out

    def robust_test(self, mode):
        out = np.zeros(4, np.uint32)
        self._check(lib().rin_robust_test(self._h, mode, out.ctypes.data))
        return {"type1": int(out[0]), "type2": int(out[1]), "type3": int(out[2]), "tested": int(out[3])}

    def set_stage_timing(self, on):
        self._check(lib().rin_set_stage_timing(self._h, 1 if on else 0))

    def launch_count(self):
        return lib().rin_get_launch_count(self._h)

    def kernel_times(self):
        e, f, t = C.c_float(), C.c_float(), C.c_float()
        self._check(lib().rin_get_kernel_times(self._h, C.byref(e), C.byref(f), C.byref(t)))
        return {"eval_ms": e.value, "filter_ms": f.value, "total_ms": t.value}

    def stage_times(self):
        k = lib().rin_num_stages()
        ms = np.zeros(k, np.float32)
        self._check(lib().rin_get_stage_times(self._h, ms.ctypes.data, k))
        return {lib().rin_stage_name(i).decode(): float(ms[i]) for i in range(k)}
