import sys, time
import os; R_=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R_+'/tests'); sys.path.insert(0,R_+'/robust-implicit-surface-networks_b200')
import numpy as np
from helpers import *
import rin_b200 as rin
from gpu_compare import compare_ia
ctx = rin.Context(0)
for R,cfg in [(16,'C2'),(32,'C2'),(24,'C4'),(64,'C2')]:
    pts,tets=orc_grid(R)
    funcs=make_funcs(synthetic_functions(cfg))
    vals=orc_eval(funcs,pts)
    port=orc_run('ia',pts,tets,vals)
    assert port.error=='', port.error
    # path A: generated grid + parametric functions on device
    ctx.generate_grid(R); ctx.set_functions(funcs)
    t=time.time(); cnt=ctx.run(); dt=time.time()-t
    gp,gt=ctx.download_grid(len(pts),len(tets))
    assert np.array_equal(gp,pts), 'grid pts'
    assert np.array_equal(gt.astype(np.uint64),tets), 'grid tets'
    gv=ctx.download_values()
    assert np.array_equal(gv,vals), ('vals', np.abs(gv-vals).max())
    mesh=ctx.download_mesh()
    compare_ia(ctx,mesh,port,cnt)
    print(R,cfg,'generated OK',cnt.as_dict(), '%.3fs'%dt, ctx.stage_times())
    # path B: host arrays
    ctx.set_mesh(pts,tets); ctx.set_values(vals)
    cnt=ctx.run(); mesh=ctx.download_mesh(); compare_ia(ctx,mesh,port,cnt)
    print(R,cfg,'host-arrays OK')
    for flags in (0, rin.FLAG_LOOKUP):
        port2=orc_run('ia',pts,tets,vals,flags=flags)
        cnt=ctx.run(flags=flags); mesh=ctx.download_mesh(); compare_ia(ctx,mesh,port2,cnt)
        print(R,cfg,'flags',flags,'OK general tets',cnt.num_general_tets)
print('ALL OK')
