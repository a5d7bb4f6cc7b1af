from _stub_common import module_getattr, _Missing
ast = _Missing("openqasm3.ast")
parser = _Missing("openqasm3.parser")
__getattr__ = module_getattr("openqasm3")
