"""Measure the FP64 / FP32 / FP32x2 FMA pipe peaks on this GPU (roofline denominators)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from integrated_path_planning_b200 import _lib

lib = _lib.load()
out = {}
for kind, name in ((0, "fp64_fma_tflops"), (1, "fp32_fma_tflops"), (2, "fp32x2_fma_tflops")):
    v = C.c_double()
    _lib.check(lib.fot_probe_fma_tflops(0, kind, C.byref(v)), "probe")
    out[name] = round(v.value, 2)
print(json.dumps(out))
