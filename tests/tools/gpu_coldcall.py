"""First call of a cold process through bn_main_fun (the R drop-in call): where the time goes."""
import os, sys, time
t_imp = time.perf_counter()
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
z = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "network_p3sim8.npz"))
from bayesnetworks_b200 import main_fun, _lib
L = _lib.lib()
t0 = time.perf_counter()
if os.environ.get("PREINIT") == "1":
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12") if False else None
n = L.bn_device_count()
t1 = time.perf_counter()
print(f"bn_device_count (driver init): {1e3 * (t1 - t0):.1f} ms", flush=True)
for i in range(3):
    t0 = time.perf_counter()
    r = main_fun(z["X"], z["source"], z["target"], np.arange(81, dtype=np.int32), z["node_type"], MaxPar=int(os.environ.get("MP", 50)),
                 N=50000, output=100, rng="rmt", seed=1234)
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
