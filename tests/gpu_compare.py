"""Shared comparison of a GPU run against an oracle run (port or hybrid reference)."""
import numpy as np


def compare_ia(ctx, mesh, port, counts, check_verts=True, xyz_tol=0.0):
    st = port["stats"]
    names = ["num_pts", "num_tets", "num_degenerate_vertex", "num_intersecting_tet", "num_k1", "num_k2",
             "num_kmore", "num_verts", "num_faces"]
    got = [getattr(counts, n) for n in names]
    assert got == st.tolist(), (got, st.tolist())
    # active sets
    fit, start = ctx.download_active()
    assert np.array_equal(start.astype(np.int64), port["start_index_of_tet"])
    assert np.array_equal(fit.astype(np.int64), port["func_in_tet"])
    # faces
    assert np.array_equal(mesh["face_offsets"].astype(np.int64), port["face_offsets"])
    assert np.array_equal(mesh["face_verts"].astype(np.int64), port["face_verts"])
    assert np.array_equal(mesh["face_tet_offsets"].astype(np.int64), port["face_tet_offsets"])
    assert np.array_equal(mesh["face_tets"].astype(np.int64).ravel(), port["face_tets"])
    ff = mesh["face_funcs"].astype(np.int64)
    ff[ff == 0xFFFFFFFF] = -1
    assert np.array_equal(ff.ravel(), port["face_funcs"])
    if check_verts:
        rec = port["vert_rec"].reshape(-1, 10)
        assert np.array_equal(mesh["vert_tet"].astype(np.int64), rec[:, 0])
        assert np.array_equal(mesh["vert_local"].astype(np.int64), rec[:, 1])
        assert np.array_equal(mesh["vert_simplex_size"].astype(np.int64), rec[:, 2])
        sv = mesh["vert_simplex"].astype(np.int64)
        sv[sv == 0xFFFFFFFF] = -1
        assert np.array_equal(sv, rec[:, 3:7])
        fi = mesh["vert_funcs"].astype(np.int64)[:, :3]
        fi[fi == 0xFFFFFFFF] = -1
        assert np.array_equal(fi, rec[:, 7:10])
    xyz = port["vert_xyz"].reshape(-1, 3)
    if xyz_tol == 0.0:
        assert np.array_equal(mesh["vert_xyz"], xyz)
    else:
        assert np.allclose(mesh["vert_xyz"], xyz, rtol=xyz_tol, atol=0)


def compare_mi(ctx, mesh, port, counts):
    st = port["stats"]
    names = ["num_pts", "num_tets", "num_degenerate_vertex", "num_intersecting_tet", "num_k1", "num_k2",
             "num_kmore", "num_verts", "num_faces"]
    got = [getattr(counts, n) for n in names]
    assert got == st.tolist(), (got, st.tolist())
    fit, start = ctx.download_active()
    assert np.array_equal(start.astype(np.int64), port["start_index_of_tet"])
    assert np.array_equal(fit.astype(np.int64), port["func_in_tet"])
    assert np.array_equal(mesh["face_offsets"].astype(np.int64), port["face_offsets"])
    assert np.array_equal(mesh["face_verts"].astype(np.int64), port["face_verts"])
    assert np.array_equal(mesh["face_tet_offsets"].astype(np.int64), port["face_tet_offsets"])
    assert np.array_equal(mesh["face_tets"].astype(np.int64).ravel(), port["face_tets"])
    ff = mesh["face_funcs"].astype(np.int64)
    ff[ff == 0xFFFFFFFF] = -1
    assert np.array_equal(ff.ravel(), port["face_funcs"])
    rec = port["vert_rec"].reshape(-1, 11)
    assert np.array_equal(mesh["vert_tet"].astype(np.int64), rec[:, 0])
    assert np.array_equal(mesh["vert_local"].astype(np.int64), rec[:, 1])
    assert np.array_equal(mesh["vert_simplex_size"].astype(np.int64), rec[:, 2])
    sv = mesh["vert_simplex"].astype(np.int64)
    sv[sv == 0xFFFFFFFF] = -1
    assert np.array_equal(sv, rec[:, 3:7])
    mi = mesh["vert_funcs"].astype(np.int64)
    mi[mi == 0xFFFFFFFF] = -1
    assert np.array_equal(mi, rec[:, 7:11])
    assert np.array_equal(mesh["vert_xyz"], port["vert_xyz"].reshape(-1, 3))
