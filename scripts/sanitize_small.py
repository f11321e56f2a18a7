"""Small IA + MI passes for compute-sanitizer (memcheck / initcheck); no oracle involved."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust-implicit-surface-networks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rin_b200 as rin
from helpers import make_funcs, synthetic_functions

for cfg, mode, R in (("C2", rin.MODE_IA, 12), ("C3", rin.MODE_MI, 12), ("C4", rin.MODE_IA, 20)):
    ctx = rin.Context(0)
    ctx.generate_grid(R)
    ctx.set_functions(make_funcs(synthetic_functions(cfg)))
    cnt = ctx.run(mode)
    mesh = ctx.download_mesh()
    act = ctx.download_active()
    print(cfg, cnt.num_intersecting_tet, cnt.num_verts, cnt.num_faces, len(mesh["vert_xyz"]), flush=True)
    ctx.close()
print("done")
