import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import bench
from noisycircuits_b200.program import PackedCircuit, Program
from noisycircuits_b200 import _lib
sq,tq,meas,pairs,instr = bench.workload()
prog=Program(PackedCircuit(bench.NQ, instr, sq, tq), independent_trajectories=True)
T=592
U=np.full((T,len(instr)),0.5)
for name,kw in (("no events",dict(rng=_lib.RNG_EXPLICIT,uniforms=U)),("philox",dict(seed=5))):
    prog.run_mcwf(T,**kw)
    best=1e9
    for r in range(3):
        t=time.perf_counter(); p,st=prog.run_mcwf(T,**kw); best=min(best,time.perf_counter()-t)
    print(name, 'traj/s %.1f'%(T/best), 'sweeps/traj %.1f'%(st['sweeps']/T), 'events/traj %.2f'%(st['events']/T), 'per-sweep us %.2f'%(best/st['sweeps']*1e6))
