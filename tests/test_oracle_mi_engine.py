"""CPU tests of the oracle's material-interface engine (restated compute_material_interface):
validated against exact rational arithmetic — every cell is the region where its material is
maximal, faces separate exactly the two materials they name, loops are oriented, the complex is a
subdivided ball."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

from helpers import oracle_lib

NONE = 0xFFFFFFFF


def mi(mats, lookup=False):
    lib = oracle_lib()
    mats = np.ascontiguousarray(mats, np.float64)
    buf = np.zeros(1 << 16, np.uint32)
    n = C.c_uint64()
    rc = lib.orc_compute_material_interface(mats.ctypes.data, len(mats), int(lookup), buf.ctypes.data, len(buf),
                                            C.byref(n))
    assert rc == 0
    w = buf[:n.value].tolist()
    nv, nf, nc, nu = w[:4]
    pos = 4
    verts = [tuple(w[pos + 4 * i:pos + 4 * i + 4]) for i in range(nv)]
    pos += 4 * nv
    faces = []
    for _ in range(nf):
        pl, nl, ln = w[pos:pos + 3]
        faces.append({"pos": pl, "neg": nl, "verts": w[pos + 3:pos + 3 + ln]})
        pos += 3 + ln
    cells = []
    for _ in range(nc):
        lab, ln = w[pos], w[pos + 1]
        cells.append({"mat": lab, "faces": w[pos + 2:pos + 2 + ln]})
        pos += 2 + ln
    uniq = w[pos + 1:pos + 1 + w[pos]] if nu else None
    return verts, faces, cells, uniq


def solve(A, b):
    n = len(A)
    M = [row[:] + [bb] for row, bb in zip(A, b)]
    for c in range(n):
        p = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        for r in range(n):
            if r != c and M[r][c] != 0:
                f = M[r][c] / M[c][c]
                M[r] = [x - f * y for x, y in zip(M[r], M[c])]
    return [M[i][n] / M[i][i] for i in range(n)]


def vertex_point(v, mats):
    """Barycentric coordinates of the point where the real materials of v tie (exact)."""
    A, rhs = [[Fraction(1)] * 4 + [Fraction(0)]], [Fraction(1)]
    for m in v:
        if m < 4:
            A.append([Fraction(int(k == m)) for k in range(4)] + [Fraction(0)])
        else:
            A.append([Fraction(float(x)) for x in mats[m - 4]] + [Fraction(-1)])
        rhs.append(Fraction(0))
    sol = solve(A, rhs)
    return sol[:4]


TET = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]


def xyz(b):
    return tuple(sum(b[i] * TET[i][k] for i in range(4)) for k in range(3))


def value(m, b, mats):
    return sum(Fraction(float(c)) * x for c, x in zip(mats[m - 4], b))


def grad_xyz(coeff):
    return (coeff[1] - coeff[0], coeff[2] - coeff[0], coeff[3] - coeff[0])


def check(mats, lookup=False):
    mats = np.asarray(mats, np.float64)
    verts, faces, cells, uniq = mi(mats, lookup)
    pts = [vertex_point(v, mats) for v in verts]
    nm = len(mats)

    def vmax(b):
        return max(value(4 + j, b, mats) for j in range(nm))

    for v, b in zip(verts, pts):
        assert all(x >= 0 for x in b)
        top = vmax(b)
        for m in v:
            if m >= 4:
                assert value(m, b, mats) == top, "vertex material does not attain the maximum"
            else:
                assert b[m] == 0
    edges = set()
    labels = set()
    for f in faces:
        n = len(f["verts"])
        assert n >= 3 and len(set(f["verts"])) == n
        for k in range(n):
            a, c = f["verts"][k], f["verts"][(k + 1) % n]
            edges.add((min(a, c), max(a, c)))
        assert f["neg"] >= 4
        key = (f["pos"], f["neg"])
        assert key not in labels, "two faces with the same label pair (not merged)"
        labels.add(key)
        P = [xyz(pts[v]) for v in f["verts"]]
        nx = ny = nz = Fraction(0)
        for k in range(n):
            (x0, y0, z0), (x1, y1, z1) = P[k], P[(k + 1) % n]
            nx += (y0 - y1) * (z0 + z1)
            ny += (z0 - z1) * (x0 + x1)
            nz += (x0 - x1) * (y0 + y1)
        if f["pos"] < 4:  # simplex boundary: CCW seen from outside
            unit = [Fraction(int(k == f["pos"])) for k in range(4)]
            g = grad_xyz(unit)
            assert nx * g[0] + ny * g[1] + nz * g[2] < 0
            for v in f["verts"]:
                assert pts[v][f["pos"]] == 0 and value(f["neg"], pts[v], mats) == vmax(pts[v])
        else:  # interface: CCW seen from the positive material's side, later material positive
            assert f["pos"] > f["neg"]
            d = [Fraction(float(a)) - Fraction(float(b)) for a, b in zip(mats[f["pos"] - 4], mats[f["neg"] - 4])]
            g = grad_xyz(d)
            assert nx * g[0] + ny * g[1] + nz * g[2] > 0
            for v in f["verts"]:
                top = vmax(pts[v])
                assert value(f["pos"], pts[v], mats) == top and value(f["neg"], pts[v], mats) == top
    assert len(verts) - len(edges) + len(faces) - len(cells) == 1
    mats_seen = set()
    for ci, c in enumerate(cells):
        assert c["mat"] not in mats_seen
        mats_seen.add(c["mat"])
        for fi in c["faces"]:
            assert c["mat"] in (faces[fi]["pos"], faces[fi]["neg"])
            for v in faces[fi]["verts"]:
                assert value(c["mat"], pts[v], mats) == vmax(pts[v])
    for fi, f in enumerate(faces):
        owners = [ci for ci, c in enumerate(cells) if fi in c["faces"]]
        assert len(owners) == (1 if f["pos"] < 4 else 2)
    # every material that is the strict maximum at some sample point owns a cell
    rng = np.random.default_rng(1)
    for _ in range(40):
        b = rng.dirichlet(np.ones(4))
        vals = mats @ b
        order = np.argsort(vals)
        if len(vals) == 1 or vals[order[-1]] - vals[order[-2]] > 1e-9:
            assert 4 + int(order[-1]) in mats_seen
    return verts, faces, cells, uniq


def test_generic_material_interfaces():
    rng = np.random.default_rng(3)
    for k in (1, 2, 3, 4, 5, 7):
        for _ in range(30):
            check(rng.uniform(-1, 1, (k, 4)))


def test_degenerate_material_interfaces():
    cases = [
        [[0, 0, 0, 0], [1, -1, -1, -1], [-1, 1, -1, -1], [-1, -1, 1, -1]],  # triple ties on edge mid points
        [[0, 0, 0, 0], [0, 1, -1, -1]],          # tie at a corner
        [[0, 0, 0, 0], [0, 0, 1, -1]],           # tie along an edge
        [[0, 0, 0, 0], [0, 0, 0, -1]],           # tie on a whole face, second never wins
        [[0, 0, 0, 0], [0, 0, 0, 1]],            # tie on a whole face, second wins elsewhere
        [[1, 2, 3, 4], [1, 2, 3, 4]],            # duplicate
        [[1, 2, 3, 4], [4, 3, 2, 1], [1, 2, 3, 4]],
        [[0, 0, 0, 0], [1, -1, 0, 0], [-1, 1, 0, 0]],
        [[0, 0, 0, 0], [1, -1, 1, -1], [1, -1, 1, -1], [-1, 1, 1, -1]],
    ]
    for m in cases:
        check(m)
    v, f, c, uniq = mi(np.array([[1, 2, 3, 4], [1, 2, 3, 4]], np.float64))
    assert uniq is not None and uniq[4] == uniq[5] and len(c) == 1


def test_two_material_table_matches_general():
    rng = np.random.default_rng(5)
    for _ in range(200):
        m = rng.uniform(-1, 1, (2, 4))
        assert mi(m, lookup=True) == mi(m, lookup=False)


def test_insertion_order_invariance_of_counts():
    rng = np.random.default_rng(9)
    for _ in range(60):
        m = rng.uniform(-1, 1, (4, 4))
        a, b = mi(m), mi(m[::-1].copy())
        assert (len(a[0]), len(a[1]), len(a[2])) == (len(b[0]), len(b[1]), len(b[2]))
