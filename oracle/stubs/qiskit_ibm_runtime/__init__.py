from _stub_common import module_getattr
__getattr__ = module_getattr("qiskit_ibm_runtime")
