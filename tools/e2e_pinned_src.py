import os, sys, time, ctypes as C, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tests import scenarios
from integrated_path_planning_b200 import BatchFrenetPlanner
Q = 4096
spline, frenet, dyn = bench.make_queries(0, Q)
pl = BatchFrenetPlanner(spline, **scenarios.S1_KNOBS)
torch.cuda.init()
rt = C.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))[0])
p = C.c_void_p()
assert rt.cudaMallocHost(C.byref(p), C.c_size_t(dyn.nbytes)) == 0
buf = np.ctypeslib.as_array((C.c_double * dyn.size).from_address(p.value)).reshape(dyn.shape)
buf[...] = dyn
for src, name in ((buf, "cudaMallocHost"), (torch.from_numpy(dyn).pin_memory().numpy(), "torch pinned"), (dyn, "pageable")):
    for _ in range(3): pl.plan_batch(frenet, 6.0, dynamic_obstacles=src[:, 0])
    t0 = time.perf_counter()
    for _ in range(8): pl.plan_batch(frenet, 6.0, dynamic_obstacles=src[:, 0])
    print(name, (time.perf_counter() - t0) / 8 * 1e3, "ms")
