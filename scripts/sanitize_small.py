"""Small passes for compute-sanitizer (memcheck / initcheck / racecheck); no oracle involved.
Covers the implicit-arrangement pipeline (tables + all general tiers), the material-interface pass (fused evaluation,
2- / 3-material tables, tabulated 3-material start for >= 4 materials, tie faces), repeated passes with learnt buffer
sizes, caller-provided meshes, mesh edges, cell-grouping maps, and a sharded run with ghost tets on a degenerate input
through the host-mediated exchange protocol (two contexts on one GPU)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust-implicit-surface-networks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rin_b200 as rin
import sharding
from helpers import make_funcs, splitmix64, synthetic_functions

for cfg, mode, R in (("C2", rin.MODE_IA, 12), ("C3", rin.MODE_MI, 12), ("C4", rin.MODE_IA, 20)):
    ctx = rin.Context(0)
    ctx.generate_grid(R)
    ctx.set_functions(make_funcs(synthetic_functions(cfg)))
    for rep in range(2):  # the second pass runs with the learnt buffer sizes (one synchronisation)
        cnt = ctx.run(mode)
    mesh = ctx.download_mesh()
    act = ctx.download_active()
    ne = ctx.mesh_edges() if hasattr(ctx, "mesh_edges") else None
    print(cfg, cnt.num_intersecting_tet, cnt.num_verts, cnt.num_faces, len(mesh["vert_xyz"]), flush=True)
    ctx.close()

# material interface with many overlapping materials: tets with >= 4 materials start from the tabulated complexes
g = splitmix64(11)
specs = [{"type": "sphere", "center": [0.5 * next(g) - 0.25 for _ in range(3)], "radius": 0.5 + 0.2 * next(g)}
         for _ in range(7)]
ctx = rin.Context(0)
ctx.generate_grid(14)
ctx.set_functions(make_funcs(specs))
cnt = ctx.run(rin.MODE_MI)
print("MI7", cnt.num_intersecting_tet, cnt.num_kmore, cnt.num_general_tets, cnt.num_verts, flush=True)
# caller-provided mesh (index-streaming filter) + values from the host
gp, gt = ctx.download_grid(15 ** 3, 5 * 14 ** 3)
vals = ctx.download_values()
ctx.set_mesh(gp, gt)
ctx.set_values(vals[:, :7].copy())
cnt = ctx.run(rin.MODE_MI)
print("MI7 host mesh", cnt.num_intersecting_tet, cnt.num_verts, flush=True)
ctx.close()

# sharded run on a degenerate input (plane x = 0 on the cut): ghost tets, host-mediated exchange, two contexts
R = 8
funcs = make_funcs([{"type": "plane", "point": [0, 0, 0], "normal": [1, 0, 0]},
                    {"type": "sphere", "center": [0, 0, 0], "radius": 0.5, "squared": True}])
ctxs = []
for rank in range(2):
    c = rin.Context(0)
    c.generate_grid(R)
    c.set_functions(funcs)
    c.set_tet_range(*sharding.slab_range(R, rank, 2))
    c.set_ghost_tets(5 * R * R, 5 * R * R)
    c.run(rin.MODE_IA)
    ctxs.append(c)
keys = [c.boundary_export(False, *c.vertex_range())[0] for c in ctxs]
n_own = [ctxs[0].mark_foreign(np.zeros((0, 4), np.uint32)), ctxs[1].mark_foreign(keys[0])]
k0, i0 = ctxs[0].boundary_export(True, *ctxs[0].vertex_range())
ctxs[0].finalize_sharded(0, np.zeros((0, 4), np.uint32), np.zeros(0, np.uint32))
ctxs[1].finalize_sharded(n_own[0], k0, i0)
print("ghost", n_own, [c.counts().num_faces for c in ctxs], flush=True)
for c in ctxs:
    c.download_mesh()
    c.close()
print("done")
