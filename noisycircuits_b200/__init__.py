"""
noisycircuits_b200 -- B200-native execution engine for the MCWF trajectory path and the density-matrix
validation path of NoisyCircuits, as a drop-in solver back-end.

    from noisycircuits_b200 import RemoteExecutor, DensityMatrixSolver, PureStateSolver   # the solver triple
    from noisycircuits_b200 import simulator                                                # FFI-level mirror
    noisycircuits_b200.install()   # registers sim_backend="b200" with an importable NoisyCircuits package

The compute path is hand-written sm_100a CUDA behind a C ABI (include/ncb200.h, libncb200.so); there is no CPU
fallback.
"""
from .solvers import DensityMatrixSolver, PureStateSolver, RemoteExecutor  # noqa: F401
from . import simulator  # noqa: F401
from .backend import install, execute, execute_sharded  # noqa: F401

__all__ = ["RemoteExecutor", "DensityMatrixSolver", "PureStateSolver", "simulator", "install", "execute", "execute_sharded"]
