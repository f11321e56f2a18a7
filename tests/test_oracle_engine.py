"""CPU tests of the oracle's per-tet engine (restated simplicial_arrangement, oracle/sa).

The library itself is un-vendored, so these tests pin the restatement by independent means:
exact rational arithmetic (fractions.Fraction) for the predicates and geometric / topological
invariants for the complexes.
"""
import ctypes as C
import itertools
from fractions import Fraction

import numpy as np
import pytest

from helpers import oracle_lib


def frac_det(m):
    n = len(m)
    if n == 1:
        return m[0][0]
    tot = Fraction(0)
    for c in range(n):
        sub = [[m[r][k] for k in range(n) if k != c] for r in range(1, n)]
        tot += (-1) ** c * m[0][c] * frac_det(sub)
    return tot


def sgn(x):
    return (x > 0) - (x < 0)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_det_sign_matches_rational_arithmetic(n):
    lib = oracle_lib()
    rng = np.random.default_rng(100 + n)
    for it in range(400):
        m = rng.uniform(-1, 1, (n, n))
        kind = it % 4
        if kind == 1:  # nearly singular: last row = combination of the others + tiny noise
            w = rng.uniform(-1, 1, n - 1)
            m[-1] = w @ m[:-1] + rng.uniform(-1, 1, n) * 1e-17
        elif kind == 2:  # exactly singular in floating point (duplicate row)
            m[-1] = m[0]
        elif kind == 3:  # small integers: many exact zeros
            m = rng.integers(-2, 3, (n, n)).astype(np.float64)
        exact = sgn(frac_det([[Fraction(float(x)) for x in row] for row in m]))
        m = np.ascontiguousarray(m)
        assert lib.orc_det_sign(n, m.ctypes.data) == exact


def arrangement(planes, lookup=False):
    lib = oracle_lib()
    planes = np.ascontiguousarray(planes, np.float64)
    buf = np.zeros(1 << 16, np.uint32)
    n = C.c_uint64()
    rc = lib.orc_compute_arrangement(planes.ctypes.data, len(planes), int(lookup), buf.ctypes.data, len(buf),
                                     C.byref(n))
    assert rc == 0
    w = buf[:n.value].tolist()
    nv, nf, nc, nu = w[:4]
    pos = 4
    verts = [tuple(w[pos + 3 * i:pos + 3 * i + 3]) for i in range(nv)]
    pos += 3 * nv
    faces = []
    for _ in range(nf):
        sp, pc, ncell, ln = w[pos:pos + 4]
        faces.append({"plane": sp, "pos": pc, "neg": ncell, "verts": w[pos + 4:pos + 4 + ln]})
        pos += 4 + ln
    cells = []
    for _ in range(nc):
        ln = w[pos]
        cells.append(w[pos + 1:pos + 1 + ln])
        pos += 1 + ln
    uniq = None
    if nu:
        npl = w[pos]
        uniq = (w[pos + 1:pos + 1 + npl], w[pos + 1 + npl:pos + 1 + 2 * npl])
    return verts, faces, cells, uniq


UNIT = [[Fraction(int(i == j)) for j in range(4)] for i in range(4)]


def vertex_bary(v, planes):
    """Exact barycentric coordinates of the intersection of three planes (ids as in the complex)."""
    rows = [UNIT[p] if p < 4 else [Fraction(float(x)) for x in planes[p - 4]] for p in v]
    rows.append([Fraction(1)] * 4)
    d = frac_det(rows)
    assert d != 0
    b = []
    for c in range(4):
        m = [r[:] for r in rows]
        for r in range(3):
            m[r][c] = Fraction(0)
        m[3] = [Fraction(int(k == c)) for k in range(4)]
        # Cramer: replace column c of A by e_4  <=> cofactor of the last row
        mm = [r[:] for r in rows]
        for r in range(4):
            mm[r][c] = Fraction(int(r == 3))
        b.append(frac_det(mm) / d)
    return b


# positively oriented reference tet for the geometric checks
TET = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]


def to_xyz(b):
    return tuple(sum(b[i] * TET[i][k] for i in range(4)) for k in range(3))


def plane_value(p, b, planes):
    coeff = UNIT[p] if p < 4 else [Fraction(float(x)) for x in planes[p - 4]]
    return sum(c * x for c, x in zip(coeff, b))


def check_complex(planes, lookup=False):
    verts, faces, cells, uniq = arrangement(planes, lookup)
    bary = [vertex_bary(v, planes) for v in verts]
    for v, b in zip(verts, bary):
        assert all(x >= 0 for x in b), "vertex outside the simplex"
        for p in v:
            assert plane_value(p, b, planes) == 0
    # Euler: V - E + F - C = 1 for a subdivided ball
    edges = set()
    for f in faces:
        n = len(f["verts"])
        assert n >= 3 and len(set(f["verts"])) == n
        for k in range(n):
            a, c = f["verts"][k], f["verts"][(k + 1) % n]
            edges.add((min(a, c), max(a, c)))
    assert len(verts) - len(edges) + len(faces) - len(cells) == 1
    # face <-> cell incidence and sides
    for fi, f in enumerate(faces):
        assert f["pos"] != 0xFFFFFFFF and fi in cells[f["pos"]]
        if f["plane"] < 4 and not (uniq and False):
            pass
        if f["neg"] != 0xFFFFFFFF:
            assert fi in cells[f["neg"]]
        for vi in f["verts"]:
            assert plane_value(f["plane"], bary[vi], planes) == 0
        # loop orientation: CCW seen from the positive side of the supporting plane
        pts = [to_xyz(bary[vi]) for vi in f["verts"]]
        nx = ny = nz = Fraction(0)
        for k in range(len(pts)):  # Newell normal
            (x0, y0, z0), (x1, y1, z1) = pts[k], pts[(k + 1) % len(pts)]
            nx += (y0 - y1) * (z0 + z1)
            ny += (z0 - z1) * (x0 + x1)
            nz += (x0 - x1) * (y0 + y1)
        coeff = UNIT[f["plane"]] if f["plane"] < 4 else [Fraction(float(x)) for x in planes[f["plane"] - 4]]
        # gradient of sum_i coeff_i b_i in xyz for TET: b0 = 1-x-y-z, b1 = x, b2 = y, b3 = z
        g = (coeff[1] - coeff[0], coeff[2] - coeff[0], coeff[3] - coeff[0])
        assert nx * g[0] + ny * g[1] + nz * g[2] > 0, "face loop is not CCW w.r.t. its plane"
    # every cell lies on the side of each of its faces that the face records
    for ci, cell in enumerate(cells):
        cv = set()
        for fi in cell:
            cv.update(faces[fi]["verts"])
        cen = [sum(bary[v][k] for v in cv) / len(cv) for k in range(4)]
        for fi in cell:
            s = sgn(plane_value(faces[fi]["plane"], cen, planes))
            assert s != 0
            assert (faces[fi]["pos"] == ci) == (s > 0)
            assert (faces[fi]["neg"] == ci) == (s < 0)
    return verts, faces, cells, uniq


def test_generic_arrangements_are_valid_complexes():
    rng = np.random.default_rng(7)
    for k in (1, 2, 3, 4, 6):
        for _ in range(25):
            check_complex(rng.uniform(-1, 1, (k, 4)))


def test_degenerate_arrangements():
    # through a vertex, an edge, a face; duplicated; opposite duplicate; three planes through a line
    cases = [
        [[0, 1, -1, -1]], [[0, 0, 1, -1]], [[0, 0, 0, -1]], [[0, 0, 0, 1]],
        [[1, 1, -1, -1], [2, 2, -2, -2]], [[1, 1, -1, -1], [-1, -1, 1, 1]],
        [[1, -1, 0, 0], [1, -1, 1, -1], [2, -2, 1, -1]],
        [[1, -1, 0, 0], [0, 0, 1, -1], [1, -1, 1, -1], [1, -1, -1, 1]],
        [[1, -1, 1, -1], [1, -1, 1, -1], [1, 1, -1, -1]],
    ]
    for planes in cases:
        verts, faces, cells, uniq = check_complex(np.array(planes, np.float64))
    # duplicate planes are reported as one group with orientation flags
    verts, faces, cells, uniq = arrangement(np.array([[1, 1, -1, -1], [-2, -2, 2, 2]], np.float64))
    groups, orient = uniq
    assert groups[4] == groups[5] and orient[4] == 1 and orient[5] == 0
    assert len(cells) == 2
    # a plane coincident with simplex face 3 joins that face's group
    verts, faces, cells, uniq = arrangement(np.array([[0, 0, 0, -3]], np.float64))
    groups, orient = uniq
    assert groups[4] == groups[3] and orient[4] == 0 and len(cells) == 1


def test_lookup_tables_agree_with_general_algorithm():
    rng = np.random.default_rng(11)
    for k in (1, 2):
        for _ in range(300):
            p = rng.uniform(-1, 1, (k, 4))
            assert arrangement(p, lookup=True) == arrangement(p, lookup=False)


def test_plane_order_invariance_of_counts():
    """The reference's robustness test (-R, src/implicit_arrangement.cpp:137-243) compares sizes
    of the forward and reversed insertion orders."""
    rng = np.random.default_rng(13)
    for _ in range(100):
        p = rng.uniform(-1, 1, (4, 4))
        a = arrangement(p)
        b = arrangement(p[::-1].copy())
        assert (len(a[0]), len(a[1]), len(a[2])) == (len(b[0]), len(b[1]), len(b[2]))
