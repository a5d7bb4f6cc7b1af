import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import random_noisy_circuit
from noisycircuits_b200.program import PackedCircuit, Program
from noisycircuits_b200 import _lib
from oracle import oracle
for n in (22, 25):
    rng=np.random.default_rng(n)
    instr,sq,tq=random_noisy_circuit(rng,n,length=40,lean=True)
    prog=Program(PackedCircuit(n,instr,sq,tq))
    print(prog.describe().split('\n')[0][:200])
    U=rng.random((2,len(instr))); U[1]=1-0.03*rng.random(len(instr))**2
    t=time.perf_counter(); probs,states,st=prog.run_mcwf(2,uniforms=U,return_states=True,rng=_lib.RNG_EXPLICIT); dt=time.perf_counter()-t
    order=prog.order()
    O=oracle.Program(n,[instr[g] for g in order],sq,tq)
    ref=O.trajectory(U[1][order],mask_fix=True)
    print(n,'replay err',np.abs(states[1]-ref).max(),'events',st['events'],'time',round(dt,2))
    t=time.perf_counter(); p,st=prog.run_mcwf(40,seed=3); dt=time.perf_counter()-t
    print(n,'40 traj norm',p.sum(),'sweeps',st['sweeps'],'time',round(dt,2))
