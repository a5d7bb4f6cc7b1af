"""TEST INFRASTRUCTURE ONLY.  Import-time stand-ins for third-party packages the reference imports at module
top level (ray, qiskit, qiskit_aer, qiskit_ibm_runtime, openqasm3) but which are not installed offline.  They
let `import NoisyCircuits` succeed so that the reference's BuildModel / Decomposition / QuantumCircuit(sim_backend=
"custom") can be run here to generate golden fixtures (tests/golden/make_fixtures.py).  Any attempt to *use* a
stubbed symbol raises."""


class _Missing:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise ImportError(f"{self._name} is a stub: the real package is not installed in this environment")

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Missing(f"{self._name}.{item}")


def module_getattr(modname):
    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Missing(f"{modname}.{name}")
    return __getattr__
