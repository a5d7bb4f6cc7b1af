import sys, numpy as np
sys.path.insert(0, '/root/repo')
from noisycircuits_b200 import _lib
from noisycircuits_b200.program import PackedCircuit, Program
from oracle import oracle
def run(n, instr, **opts):
    prog = Program(PackedCircuit(n, instr, noisy=False), _lib.MODE_PURE, **opts)
    st, _ = prog.run_pure()
    ref = oracle.Program(n, instr, noisy=False).pure_state()
    return np.abs(st - ref).max(), prog.describe().splitlines()[0]
cases = {
 'sx': [["sx",[0,0],0]],
 'sx,sx': [["sx",[0,0],0],["sx",[0,0],0]],
 'rx,rz same': [["rx",[0,0],0.3],["rz",[0,0],0.7]],
 'rx q0, rz q1': [["rx",[0,0],0.3],["rz",[1,1],0.7]],
 'rz,rz': [["rz",[0,0],0.3],["rz",[1,1],0.7]],
 'h,rz,cz': [["h",[0,0],0],["h",[1,1],0],["rz",[0,0],0.3],["cz",[0,1],0]],
 'x,sx,rz,sx': [["x",[1,1],0],["sx",[1,1],0],["rz",[1,1],-0.4],["sx",[1,1],0]],
}
for n in (2, 5, 9, 13):
    for name, instr in cases.items():
        for opts in ({}, {'fuse': 0}, {'reg_bits': 4}):
            try:
                d, desc = run(n, instr, **opts)
                print(n, name, opts, '%.2e' % d, 'OK' if d < 1e-12 else 'FAIL', desc[:90])
            except Exception as e:
                print(n, name, opts, 'EXC', e)
