"""GPU: near-degenerate inputs in the style of the reference's robustness experiment
(examples/Table2_Fig11: four spheres whose centres are 1e-7 .. 1e-6 apart; MI adds a zero
function).  The floating-point filters of the device predicates cannot decide these signs, so the
exact expansion arithmetic runs; the combinatorial output must still equal the oracle's (whose
predicates are validated against rational arithmetic)."""
import numpy as np
import pytest

from gpu_compare import compare_ia, compare_mi
from helpers import make_funcs, orc_eval, orc_grid, orc_run, splitmix64

pytestmark = pytest.mark.gpu


def near_coincident_spheres(eps, seed, with_zero=False):
    g = splitmix64(seed)
    specs = [{"type": "sphere", "center": [eps * (2 * next(g) - 1) for _ in range(3)], "radius": 0.5}
             for _ in range(4)]
    if with_zero:
        specs.append({"type": "zero"})
    return specs


@pytest.mark.parametrize("eps", [1e-7, 1e-14, 1e-15, 3e-16])
def test_ia_near_coincident_spheres(eps):
    import rin_b200 as rin
    pts, tets = orc_grid(12)
    funcs = make_funcs(near_coincident_spheres(eps, 11))
    vals = orc_eval(funcs, pts)
    port = orc_run("ia", pts, tets, vals)
    assert port.error == ""
    ctx = rin.Context(0)
    ctx.set_mesh(pts, tets)
    ctx.set_functions(funcs)
    cnt = ctx.run(rin.MODE_IA)
    compare_ia(ctx, ctx.download_mesh(), port, cnt)
    # the reference's -R self-test on the device: forward vs reversed insertion order
    r = ctx.robust_test(rin.MODE_IA)
    assert r == {"type1": 0, "type2": 0, "type3": 0, "tested": cnt.num_intersecting_tet}
    ctx.close()


@pytest.mark.parametrize("eps", [1e-7, 1e-14, 1e-15, 3e-16])
def test_mi_near_coincident_spheres(eps):
    import rin_b200 as rin
    pts, tets = orc_grid(12)
    funcs = make_funcs(near_coincident_spheres(eps, 12, with_zero=True))
    vals = orc_eval(funcs, pts)
    port = orc_run("mi", pts, tets, vals)
    assert port.error == ""
    ctx = rin.Context(0)
    ctx.set_mesh(pts, tets)
    ctx.set_functions(funcs)
    cnt = ctx.run(rin.MODE_MI)
    compare_mi(ctx, ctx.download_mesh(), port, cnt)
    r = ctx.robust_test(rin.MODE_MI)
    assert r == {"type1": 0, "type2": 0, "type3": 0, "tested": cnt.num_intersecting_tet}
    ctx.close()


def test_exact_fallback_is_exercised():
    """Planes through grid vertices: determinants are exactly zero, only exact arithmetic can tell."""
    import rin_b200 as rin
    pts, tets = orc_grid(10)
    funcs = make_funcs([{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                        {"type": "plane", "point": [0, 0, 0], "normal": [1, 1, 0]},
                        {"type": "plane", "point": [0, 0, 0], "normal": [1, 1, 1]},
                        {"type": "plane", "point": [0, 0, 0], "normal": [0, 1, -1]}])
    vals = orc_eval(funcs, pts)
    port = orc_run("ia", pts, tets, vals)
    assert port.error == ""
    ctx = rin.Context(0)
    ctx.set_mesh(pts, tets)
    ctx.set_functions(funcs)
    cnt = ctx.run(rin.MODE_IA)
    assert cnt.num_exact_fallbacks > 0
    compare_ia(ctx, ctx.download_mesh(), port, cnt)
    ctx.close()
