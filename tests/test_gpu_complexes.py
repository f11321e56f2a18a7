"""GPU: rin_get_complexes returns, for any requested tet, exactly the complex the oracle's
compute_arrangement / compute_material_interface builds (the structure the reference's host
stages read: src/pair_faces.cpp:138-238, src/topo_ray_shooting.cpp:56-57)."""
import ctypes as C

import numpy as np
import pytest

from helpers import make_funcs, oracle_lib, orc_eval, orc_grid, orc_run, synthetic_functions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,cfg,R", [("ia", "C2", 14), ("ia", "C4", 24), ("mi", "C3", 14)])
def test_complexes_match_oracle(mode, cfg, R):
    import rin_b200 as rin
    funcs = make_funcs(synthetic_functions(cfg))
    pts, tets = orc_grid(R)
    vals = orc_eval(funcs, pts)
    port = orc_run(mode, pts, tets, vals)
    ctx = rin.Context(0)
    ctx.generate_grid(R)
    ctx.set_functions(funcs)
    m = rin.MODE_IA if mode == "ia" else rin.MODE_MI
    ctx.run(m)
    k = np.diff(port["start_index_of_tet"])
    active = np.nonzero(k > 0)[0]
    inactive = np.nonzero(k == 0)[0][:5]
    req = np.concatenate([active[::3], inactive, active[:7][::-1]]).astype(np.uint64)  # any order, repeats
    off, words = ctx.get_complexes(m, req)
    lib = oracle_lib()
    buf = np.zeros(1 << 16, np.uint32)
    n = C.c_uint64()
    for i, t in enumerate(req.tolist()):
        lib.orc_get_complex(port._h, t, buf.ctypes.data, len(buf), C.byref(n))
        got = words[int(off[i]):int(off[i + 1])]
        assert np.array_equal(got, buf[:n.value]), (mode, t)
    ctx.close()
