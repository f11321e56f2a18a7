import sys, time, os
R_=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R_+'/tests'); sys.path.insert(0,R_+'/robust-implicit-surface-networks_b200')
import numpy as np
from helpers import *
import rin_b200 as rin
R=int(sys.argv[1]) if len(sys.argv)>1 else 128
cfg=sys.argv[2] if len(sys.argv)>2 else 'C2'
flags=int(sys.argv[3]) if len(sys.argv)>3 else 3
ctx = rin.Context(0)
funcs=make_funcs(synthetic_functions(cfg))
ctx.generate_grid(R); ctx.set_functions(funcs)
for it in range(5):
    t=time.time(); cnt=ctx.run(flags=flags); dt=time.time()-t
    st=ctx.stage_times()
    print('iter',it,'wall %.3f ms'%(dt*1e3), 'sum stages %.3f'%sum(st.values()), {k:round(v,3) for k,v in st.items()})
print(cnt.as_dict())
print('tets/s %.3e'%(cnt.num_tets/dt))
