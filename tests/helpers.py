"""ctypes access to the oracle libraries (test infrastructure) and shared input builders.

liboracle.so        -- CPU restatement of the hot path (oracle/port + oracle/sa)
_ref/libref_hybrid.so -- the reference's own sources compiled in place against the restated
                       per-tet engine (only present where it was built; it travels prebuilt).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

FN_PLANE, FN_SPHERE, FN_CYLINDER, FN_ZERO, FN_TORUS = 0, 1, 2, 3, 4
FLAG_LOOKUP, FLAG_SECONDARY, FLAG_NEGATE = 1, 2, 4

FUNC_DESC = np.dtype([("type", "<i4"), ("flip", "<i4"), ("p", "<f8", (10,))])


def _build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
    return so


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(_build_oracle())
        lib.orc_ia_run.restype = C.c_void_p
        lib.orc_ia_run.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                   C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]
        if hasattr(lib, "orc_mi_run"):
            lib.orc_mi_run.restype = C.c_void_p
            lib.orc_mi_run.argtypes = lib.orc_ia_run.argtypes
        for nm, rt in (("orc_i64", C.POINTER(C.c_int64)), ("orc_f64", C.POINTER(C.c_double))):
            fn = getattr(lib, nm)
            fn.restype = rt
            fn.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
        lib.orc_error.restype = C.c_char_p
        lib.orc_error.argtypes = [C.c_void_p]
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_generate_grid.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_eval_functions.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_get_complex.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                        C.POINTER(C.c_uint64)]
        lib.orc_compute_arrangement.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p,
                                                C.c_uint64, C.POINTER(C.c_uint64)]
        lib.orc_compute_material_interface.argtypes = lib.orc_compute_arrangement.argtypes
        _oracle = lib
    return _oracle


def ref_lib():
    """The hybrid reference, or None when it has not been built (it needs /root/reference)."""
    global _ref
    if _ref is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libref_hybrid.so")
        if not os.path.exists(so):
            return None
        lib = C.CDLL(so)
        for nm in ("ref_ia_run", "ref_mi_run"):
            fn = getattr(lib, nm)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                           C.c_uint32]
        for nm, rt in (("ref_i64", C.POINTER(C.c_int64)), ("ref_f64", C.POINTER(C.c_double))):
            fn = getattr(lib, nm)
            fn.restype = rt
            fn.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
        lib.ref_error.restype = C.c_char_p
        lib.ref_error.argtypes = [C.c_void_p]
        lib.ref_free.argtypes = [C.c_void_p]
        _ref = lib
    return _ref


_dropin = None


def dropin_lib():
    """Reference host stages + csg on top of the GPU library (oracle/_ref/libref_gpu_dropin.so), or None."""
    global _dropin
    if _dropin is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libref_gpu_dropin.so")
        if not os.path.exists(so):
            return None
        lib = C.CDLL(so)
        for nm in ("ref_ia_run", "ref_mi_run"):
            fn = getattr(lib, nm)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32]
        for nm, rt in (("ref_i64", C.POINTER(C.c_int64)), ("ref_f64", C.POINTER(C.c_double))):
            fn = getattr(lib, nm)
            fn.restype = rt
            fn.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
        lib.ref_error.restype = C.c_char_p
        lib.ref_error.argtypes = [C.c_void_p]
        lib.ref_free.argtypes = [C.c_void_p]
        _dropin = lib
    return _dropin


class Bag:
    """Named result vectors of one oracle / reference run (copied out, handle freed)."""

    def __init__(self, lib, prefix, handle, names_i64, names_f64):
        self.error = getattr(lib, prefix + "_error")(handle).decode()
        self.d = {}
        n = C.c_uint64()
        for nm in names_i64:
            p = getattr(lib, prefix + "_i64")(handle, nm.encode(), C.byref(n))
            self.d[nm] = np.ctypeslib.as_array(p, (n.value,)).copy() if n.value else np.zeros(0, np.int64)
        for nm in names_f64:
            p = getattr(lib, prefix + "_f64")(handle, nm.encode(), C.byref(n))
            self.d[nm] = np.ctypeslib.as_array(p, (n.value,)).copy() if n.value else np.zeros(0, np.float64)
        self._lib, self._prefix, self._h = lib, prefix, handle

    def __getitem__(self, k):
        return self.d[k]

    def close(self):
        if self._h:
            getattr(self._lib, self._prefix + "_free")(self._h)
            self._h = None

    def __del__(self):
        self.close()


MESH_I64 = ["stats", "face_offsets", "face_verts", "face_tet_offsets", "face_tets", "face_funcs"]
MAP_I64 = ["global_vId_of_tet_vert", "global_vId_start_index_of_tet", "iso_fId_of_tet_face",
           "iso_fId_start_index_of_tet"]
PORT_I64 = MESH_I64 + ["func_in_tet", "start_index_of_tet", "vert_rec", "engine"] + MAP_I64
REF_I64 = MESH_I64 + ["success", "threw", "patches", "patches_offsets", "chains", "chains_offsets",
                      "shells", "shells_offsets", "cells", "cells_offsets", "patch_function_label",
                      "cell_function_label", "edges", "edge_faces", "edge_faces_offsets", "timing_label_bytes", "stats_label_bytes",
                      "non_manifold_edges_of_vert", "non_manifold_edges_of_vert_offsets"]


def make_funcs(specs):
    """specs: list of dicts in the reference's function-file JSON schema -> FUNC_DESC array."""
    out = np.zeros(len(specs), FUNC_DESC)
    for i, s in enumerate(specs):
        t = s["type"]
        p = out[i]["p"]
        out[i]["flip"] = 1 if s.get("is_flipped", False) else 0
        if t == "plane":
            out[i]["type"] = FN_PLANE
            p[0:3] = s["point"]
            p[3:6] = s["normal"]
        elif t == "sphere":
            out[i]["type"] = FN_SPHERE
            p[0:3] = s["center"]
            p[3] = s["radius"]
            p[4] = 1.0 if s.get("squared", False) else 0.0
        elif t == "cylinder":
            out[i]["type"] = FN_CYLINDER
            p[0:3] = s["axis_point"]
            p[3:6] = s["axis_vector"]
            p[6] = s["radius"]
        elif t == "torus":
            out[i]["type"] = FN_TORUS
            p[0:3] = s["center"]
            p[3:6] = s["axis_vector"]
            p[6] = s["major_radius"]
            p[7] = s["minor_radius"]
        elif t == "zero":
            out[i]["type"] = FN_ZERO
        else:
            raise ValueError("function type %r is out of scope" % t)
    return out


def load_funcs(path):
    with open(path) as f:
        return make_funcs(json.load(f))


def orc_grid(R, bmin=(-1, -1, -1), bmax=(1, 1, 1)):
    lib = oracle_lib()
    N = R + 1
    pts = np.empty((N ** 3, 3), np.float64)
    tets = np.empty((5 * R ** 3, 4), np.uint64)
    a = np.asarray(bmin, np.float64)
    b = np.asarray(bmax, np.float64)
    rc = lib.orc_generate_grid(R, a.ctypes.data, b.ctypes.data, pts.ctypes.data, tets.ctypes.data)
    assert rc == 0
    return pts, tets


def orc_eval(funcs, pts):
    lib = oracle_lib()
    out = np.empty((len(pts), len(funcs)), np.float64)
    pts = np.ascontiguousarray(pts, np.float64)
    lib.orc_eval_functions(funcs.ctypes.data, len(funcs), pts.ctypes.data, len(pts), out.ctypes.data)
    return out


def _prep(pts, tets, vals):
    pts = np.ascontiguousarray(pts, np.float64)
    tets = np.ascontiguousarray(tets, np.uint64)
    vals = np.ascontiguousarray(vals, np.float64)
    return pts, tets, vals


def orc_run(mode, pts, tets, vals, flags=FLAG_LOOKUP | FLAG_SECONDARY, tet_first=0, tet_count=0):
    lib = oracle_lib()
    pts, tets, vals = _prep(pts, tets, vals)
    fn = lib.orc_ia_run if mode == "ia" else lib.orc_mi_run
    h = fn(pts.ctypes.data, len(pts), tets.ctypes.data, len(tets), vals.ctypes.data, vals.shape[1],
           flags, tet_first, tet_count)
    return Bag(lib, "orc", h, PORT_I64, ["vert_xyz", "timings"])


def ref_run(mode, pts, tets, vals, robust=False, lookup=True, secondary=True, ray=True, quiet=True, lib=None):
    lib = lib or ref_lib()
    assert lib is not None
    pts, tets, vals = _prep(pts, tets, vals)
    flags = (1 if robust else 0) | (2 if lookup else 0) | (4 if secondary else 0) | (8 if ray else 0) | \
        (16 if quiet else 0)
    fn = lib.ref_ia_run if mode == "ia" else lib.ref_mi_run
    h = fn(pts.ctypes.data, len(pts), tets.ctypes.data, len(tets), vals.ctypes.data, vals.shape[1], flags)
    b = Bag(lib, "ref", h, REF_I64, ["vert_xyz", "timings"])
    b.timing_labels = bytes(b["timing_label_bytes"].astype(np.uint8)).decode().split("\n")[:-1]
    b.stats_labels = bytes(b["stats_label_bytes"].astype(np.uint8)).decode().split("\n")[:-1]
    b.stats = dict(zip(b.stats_labels, b["stats"].tolist()))
    return b


def ref_cellgroup_maps(tets, vals, n_pts):
    """The reference's own second extract_iso_mesh overload (src/extract_mesh.cpp:268-566) on these inputs."""
    lib = ref_lib()
    tets = np.ascontiguousarray(tets, np.uint64)
    vals = np.ascontiguousarray(vals, np.float64)
    lib.ref_ia_cellgroup_maps.restype = C.c_void_p
    lib.ref_ia_cellgroup_maps.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32]
    h = lib.ref_ia_cellgroup_maps(tets.ctypes.data, len(tets), n_pts, vals.ctypes.data, vals.shape[1])
    return Bag(lib, "ref", h, ["global_vId_of_tet_vert", "global_vId_start_index_of_tet", "iso_fId_of_tet_face",
                               "iso_fId_start_index_of_tet", "counts", "cell_graph"], [])


def ref_mi_cellgroup_maps(tets, vals, material_in_tet, start_index_of_tet):
    """The reference's own second extract_MI_mesh overload (src/extract_mesh.cpp:988-1443)."""
    lib = ref_lib()
    tets = np.ascontiguousarray(tets, np.uint64)
    vals = np.ascontiguousarray(vals, np.float64)
    mit = np.ascontiguousarray(material_in_tet, np.int64)
    st = np.ascontiguousarray(start_index_of_tet, np.int64)
    lib.ref_mi_cellgroup_maps.restype = C.c_void_p
    lib.ref_mi_cellgroup_maps.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                          C.c_void_p]
    h = lib.ref_mi_cellgroup_maps(tets.ctypes.data, len(tets), vals.ctypes.data, vals.shape[1], mit.ctypes.data,
                                  len(mit), st.ctypes.data)
    return Bag(lib, "ref", h, ["global_vId_of_tet_vert", "global_vId_start_index_of_tet", "iso_fId_of_tet_face",
                               "iso_fId_start_index_of_tet", "counts", "cell_graph"], [])


def ref_csg(pts, tets, vals, expr, positive_inside=True, lib=None):
    """csg() of the reference with one of its test expressions (see oracle/ref_capi.cpp)."""
    lib = lib or ref_lib()
    pts, tets, vals = _prep(pts, tets, vals)
    lib.ref_csg_run.restype = C.c_void_p
    lib.ref_csg_run.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32,
                                C.c_int, C.c_int]
    h = lib.ref_csg_run(pts.ctypes.data, len(pts), tets.ctypes.data, len(tets), vals.ctypes.data, vals.shape[1],
                        2 | 4 | 8 | 16, expr, int(positive_inside))
    return Bag(lib, "ref", h, ["success", "threw", "patches", "patches_offsets", "chains", "chains_offsets",
                               "non_manifold_edges_of_vert", "non_manifold_edges_of_vert_offsets",
                               "patch_sign_label"], [])


def crs(bag, name):
    off = bag[name + "_offsets"]
    dat = bag[name]
    return [dat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def splitmix64(seed):
    """The deterministic generator of SURVEY section 8(d): yields u in [0,1)."""
    s = seed & 0xFFFFFFFFFFFFFFFF
    while True:
        s = (s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z = z ^ (z >> 31)
        yield (z >> 11) * (1.0 / 9007199254740992.0)


def synthetic_functions(config):
    """Function sets of the BASELINE configurations."""
    specs = []
    if config in ("C2", "C5"):
        g = splitmix64(1)
        for f in range(8):
            if f % 2 == 0:
                c = [next(g) - 0.5 for _ in range(3)]
                r = 0.2 + 0.4 * next(g)
                specs.append({"type": "sphere", "center": c, "radius": r, "squared": True})
            else:
                p = [next(g) - 0.5 for _ in range(3)]
                while True:
                    n = [2 * next(g) - 1 for _ in range(3)]
                    l2 = sum(x * x for x in n)
                    if 1e-6 < l2 <= 1.0:
                        break
                l = l2 ** 0.5
                specs.append({"type": "plane", "point": p, "normal": [x / l for x in n]})
    elif config == "C3":
        g = splitmix64(2)
        for f in range(6):
            c = [1.2 * next(g) - 0.6 for _ in range(3)]
            r = 0.4 + 0.4 * next(g)
            specs.append({"type": "sphere", "center": c, "radius": r, "squared": False})
    elif config == "C4":
        g = splitmix64(3)
        for f in range(32):
            c = [0.1 * next(g) - 0.05 for _ in range(3)]
            r = 0.45 + 0.1 * next(g)
            specs.append({"type": "sphere", "center": c, "radius": r, "squared": True})
    elif config == "C2cyl":
        # SURVEY 8(d) cylinder swap: C2 with the spheres 0 and 4 replaced by cylinders (seed 11)
        specs = synthetic_functions("C2")
        g = splitmix64(11)
        for f in (0, 4):
            ap = [next(g) - 0.5 for _ in range(3)]
            while True:
                n = [2 * next(g) - 1 for _ in range(3)]
                l2 = sum(x * x for x in n)
                if 1e-6 < l2 <= 1.0:
                    break
            l = l2 ** 0.5
            specs[f] = {"type": "cylinder", "axis_point": ap, "axis_vector": [x / l for x in n],
                        "radius": 0.1 + 0.3 * next(g)}
    elif config == "TOR":
        # cylinders, tori, a plane and a sphere (every parametric type of the device evaluator; seed 12)
        g = splitmix64(12)

        def unit():
            while True:
                n = [2 * next(g) - 1 for _ in range(3)]
                l2 = sum(x * x for x in n)
                if 1e-6 < l2 <= 1.0:
                    l = l2 ** 0.5
                    return [x / l for x in n]
        for f in range(3):
            specs.append({"type": "torus", "center": [0.6 * next(g) - 0.3 for _ in range(3)], "axis_vector": unit(),
                          "major_radius": 0.3 + 0.3 * next(g), "minor_radius": 0.08 + 0.1 * next(g)})
        for f in range(2):
            specs.append({"type": "cylinder", "axis_point": [next(g) - 0.5 for _ in range(3)], "axis_vector": unit(),
                          "radius": 0.1 + 0.3 * next(g), "is_flipped": bool(f)})
        specs.append({"type": "plane", "point": [0.1, -0.2, 0.05], "normal": unit()})
        specs.append({"type": "sphere", "center": [0.0, 0.1, -0.1], "radius": 0.7, "squared": False})
    else:
        raise ValueError(config)
    return specs
