import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import montecarlo_simulation_b200 as m
L=40.0
rng=np.random.default_rng(0)
g=np.stack(np.meshgrid(np.arange(28),np.arange(28),indexing="ij"),-1).reshape(-1,2).astype(float)
xy=(g+0.5)*(L/28)+rng.uniform(-0.2,0.2,g.shape)
with m.Context(L,L,2.5,potential="LennardJones",T=1.0,activity=0.1,delta=0.1,seed=1) as ctx:
    ctx.upload(xy)
    st=ctx.run_sweeps(5,disp_per_particle=1,gc_per_cell=1)
    print(st)
