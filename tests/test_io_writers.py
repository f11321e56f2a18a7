"""the product's config parser, tet-mesh loader and result writers (host/rin_io.cpp) against the
reference's own io.cpp compiled in place (oracle/_ref/libref_io.so: real nlohmann::json, std::filesystem behind
the ghc shim, the MSH serialiser behind the mshio shim).  mesh.json (IA / MI / CSG variants), timings.json,
stats.json and the three .msh files must be identical (JSON: byte for byte outside the rare doubles that
nlohmann's Grisu2 prints with a 17th digit); the .msh files are also read back per the MSH 4.1
layout (byte parity with MshIO itself is unpinned: the library is absent)."""
import ctypes as C
import json
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
FILES = ["mesh.json", "mesh_mi.json", "mesh_csg.json", "timings.json", "stats.json", "mesh_chains.msh",
         "mesh_patches.msh", "mesh_cells.msh"]


class _Lib:
    """Every call runs in a process of its own: the reference library (nlohmann::json, built with the image's
    newer g++) and the product library must not share one libstdc++ locale state inside the test process."""

    def __init__(self, path):
        self.path = path

    def _run(self, code):
        p = subprocess.run([sys.executable, "-c", "import ctypes as C, sys\nlib = C.CDLL(%r)\n%s" % (self.path, code)],
                           capture_output=True)
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        return p.stdout

    def io_write_all(self, d, seed, n_pts, n_faces):
        code = ("lib.io_write_all.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int]\n"
                "sys.exit(lib.io_write_all(%r, %d, %d, %d))" % (d, seed, n_pts, n_faces))
        self._run(code)
        return 0

    def io_read_all(self, cfg):
        code = ("lib.io_read_all.argtypes = [C.c_char_p, C.c_char_p, C.c_int]\nbuf = C.create_string_buffer(1 << 16)\n"
                "n = lib.io_read_all(%r, buf, len(buf))\nassert n > 0\nsys.stdout.buffer.write(buf.value)" % cfg)
        return self._run(code)


@pytest.fixture(scope="module")
def product():
    so = os.path.join(ORACLE, "libproduct_io.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", ORACLE, "libproduct_io.so"])
    return _Lib(so)


@pytest.fixture(scope="module")
def reference():
    so = os.path.join(ORACLE, "_ref", "libref_io.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_io.so not built (needs /root/reference)")
    return _Lib(so)


NUM = re.compile(rb"-?\d+(?:\.\d+)?(?:[eE][+-]?\d+)?")


def same_json_text(a, b):
    """Byte equality outside the numbers; numbers must denote the same value in the same notation class (integer /
    fixed / scientific).  nlohmann prints doubles with Grisu2, which now and then emits a 17th digit where 16
    round-trip (0.5845329003030451 for 0.584532900303045); the product prints the shortest digits."""
    ta, tb = NUM.split(a), NUM.split(b)
    if ta != tb:
        return False, "structure differs"
    na, nb = NUM.findall(a), NUM.findall(b)
    diff = 0
    for x, y in zip(na, nb):
        if x != y:
            if float(x) != float(y) or (b"e" in x) != (b"e" in y) or (b"." in x) != (b"." in y):
                return False, "number %r vs %r" % (x, y)
            if len(y) > len(x):
                return False, "product longer than the reference: %r vs %r" % (x, y)
            diff += 1
    return True, diff / max(1, len(na))


@pytest.mark.parametrize("seed,n_pts,n_faces", [(1, 2000, 1500), (2, 17, 9), (3, 0, 0), (4, 50000, 30000)])
def test_result_files_are_byte_identical_to_the_reference_writers(product, reference, tmp_path, seed, n_pts, n_faces):
    a, b = tmp_path / "ref", tmp_path / "ours"
    a.mkdir()
    b.mkdir()
    assert reference.io_write_all(str(a).encode(), seed, n_pts, n_faces) == 0
    assert product.io_write_all(str(b).encode(), seed, n_pts, n_faces) == 0
    for f in FILES:
        fa, fb = a / f, b / f
        assert fa.exists() == fb.exists(), f
        if fa.exists():
            if f.endswith(".json"):
                ok, info = same_json_text(fa.read_bytes(), fb.read_bytes())
                assert ok, (f, info)
                assert info < 0.01, (f, info)  # share of doubles where Grisu2 spends one digit more
            else:
                assert fa.read_bytes() == fb.read_bytes(), f
    if n_pts == 0:
        assert (b / "mesh.json").read_text() == ('{"cells":null,"cells_label":null,"chains":null,"corners":null,'
                                                 '"edges":null,"faces":null,"patches":null,"patches_label":[[]],'
                                                 '"points":null,"shells":null}\n')
        assert (b / "timings.json").read_text() == "null\n"
    else:
        json.loads((b / "mesh.json").read_text())  # well-formed


def read_msh41(path):
    """Minimal MSH 4.1 binary reader (data-size 8): nodes, elements and element data."""
    raw = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        e = raw.index(b"\n", pos)
        s = raw[pos:e].decode()
        pos = e + 1
        return s

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, raw, pos)
        pos += struct.calcsize("<" + fmt)
        return v

    assert line() == "$MeshFormat"
    assert line() == "4.1 1 8"
    assert take("i") == (1,)
    assert line() == "" and line() == "$EndMeshFormat"
    out = {"nodes": [], "elements": [], "element_data": {}}
    while pos < len(raw):
        sec = line()
        if sec == "$Nodes":
            nb, nn, lo, hi = take("4Q")
            for _ in range(nb):
                dim, tag, par, n = take("3iQ")
                tags = np.frombuffer(raw, "<u8", n, pos)
                pos += 8 * n
                xyz = np.frombuffer(raw, "<f8", 3 * n, pos).reshape(-1, 3)
                pos += 24 * n
                out["nodes"].append((dim, tag, tags, xyz))
            assert line() == "" and line() == "$EndNodes"
            assert sum(len(b[2]) for b in out["nodes"]) == nn and (lo, hi) == (1, nn)
        elif sec == "$Elements":
            nb, ne, lo, hi = take("4Q")
            for _ in range(nb):
                dim, tag, typ, n = take("3iQ")
                k = {1: 2, 2: 3, 4: 4}[typ] + 1
                data = np.frombuffer(raw, "<u8", k * n, pos).reshape(-1, k)
                pos += 8 * k * n
                out["elements"].append((dim, tag, typ, data))
            assert line() == "" and line() == "$EndElements"
            assert sum(len(b[3]) for b in out["elements"]) == ne
        elif sec == "$ElementData":
            assert line() == "1"
            name = line().strip('"')
            assert line() == "1" and line() == "0"
            assert line() == "5"
            ints = [int(line()) for _ in range(5)]
            rec = np.frombuffer(raw, np.dtype([("tag", "<i4"), ("v", "<f8")]), ints[2], pos)
            pos += 12 * ints[2]
            assert line() == "" and line() == "$EndElementData"
            out["element_data"][name] = rec
        else:
            raise AssertionError("unexpected section " + sec)
    return out


def test_msh_files_read_back(product, tmp_path):
    assert product.io_write_all(str(tmp_path).encode(), 7, 300, 200) == 0
    mesh = json.loads((tmp_path / "mesh.json").read_text())
    pts = np.array(mesh["points"])
    chains = read_msh41(tmp_path / "mesh_chains.msh")
    assert len(chains["nodes"]) == len(mesh["chains"]) == len(chains["elements"])
    for (dim, tag, tags, xyz), (edim, etag, typ, data), chain in zip(chains["nodes"], chains["elements"], mesh["chains"]):
        assert dim == edim == 1 and typ == 1 and tag == etag and len(data) == len(chain)
        for row, e in zip(data, chain):
            v1, v2 = mesh["edges"][e]
            assert np.array_equal(xyz[row[1] - tags[0]], pts[v1]) and np.array_equal(xyz[row[2] - tags[0]], pts[v2])
    patches = read_msh41(tmp_path / "mesh_patches.msh")
    n_tri = sum(len(mesh["faces"][f]) - 2 for p in mesh["patches"] for f in p)
    assert sum(len(b[3]) for b in patches["elements"]) == n_tri
    pid = patches["element_data"]["patch_id"]
    poly = patches["element_data"]["polygon_id"]
    assert len(pid) == len(poly) == n_tri and np.array_equal(pid["tag"], np.arange(1, n_tri + 1))
    expect_pid = [i for i, p in enumerate(mesh["patches"]) for f in p for _ in range(len(mesh["faces"][f]) - 2)]
    expect_poly = [f for p in mesh["patches"] for f in p for _ in range(len(mesh["faces"][f]) - 2)]
    assert pid["v"].tolist() == expect_pid and poly["v"].tolist() == expect_poly
    cells = read_msh41(tmp_path / "mesh_cells.msh")
    assert len(cells["nodes"]) == len(mesh["cells"]) and "cell_id" in cells["element_data"]


def test_config_and_tet_mesh_readers(product, reference, tmp_path):
    (tmp_path / "sub").mkdir()
    mesh = [[[0, 0, 0], [1, 0.5, 0], [0, 1e-3, 0], [0.25, 0, -1]], [[0, 1, 2, 3]]]
    (tmp_path / "sub" / "tet.json").write_text(json.dumps(mesh))
    cfgs = {
        "grid.json": {"gridResolution": 37, "gridBbox": [[-1, -1.5, -2], [1, 1.25, 2e0]], "funcFile": "f/funcs.json",
                      "outputDir": "out", "useLookup": True, "useSecondaryLookup": False, "useTopoRayShooting": True},
        "mesh.json": {"tetMeshFile": "sub/tet.json", "gridResolution": 5, "gridBbox": [[0, 0, 0], [1, 1, 1]],
                      "funcFile": "/abs/funcs.json", "outputDir": "../o", "useLookup": False,
                      "useSecondaryLookup": False, "useTopoRayShooting": False},
    }
    for name, cfg in cfgs.items():
        p = tmp_path / name
        p.write_text(json.dumps(cfg, indent=1))
        ra = reference.io_read_all(str(p).encode())
        rb = product.io_read_all(str(p).encode())
        assert ra == rb and b"exception" not in rb, (name, ra, rb)
    for lib in (reference, product):
        assert lib.io_read_all(str(tmp_path / "missing.json").encode()) == b"exception: Config file does not exist!"


def test_number_notation_matches_nlohmann_on_many_doubles(product, reference, tmp_path):
    """360 k doubles of every magnitude through both writers (the points array of mesh.json)."""
    a, b = tmp_path / "ref", tmp_path / "ours"
    a.mkdir()
    b.mkdir()
    assert reference.io_write_all(str(a).encode(), 99, 120000, 1) == 0
    assert product.io_write_all(str(b).encode(), 99, 120000, 1) == 0
    ok, info = same_json_text((a / "mesh.json").read_bytes(), (b / "mesh.json").read_bytes())
    assert ok and info < 0.01, info
