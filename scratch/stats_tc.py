import sys, os, ctypes
os.environ["VLG_B200_LIB"] = "scratch/libvlg_stats.so"
sys.path.insert(0, ".")
import numpy as np, torch
import vlg_b200, bench
from vlg_b200 import _lib
nc = 148
w, a, b, om, _ = bench.synthetic_workload(nc)
dev = "cuda"
dec = vlg_b200.DecoderEnsemble.from_arrays(*[w[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")], dev)
basis, _ = vlg_b200.construct_nullspace_basis(4)
m = vlg_b200.GeodesicSplineBatch(a.to(dev), b.to(dev), basis.to(dev), om.to(dev), 4)
t = torch.linspace(0, 1, 2000, device=dev)
for _ in range(2):
    vlg_b200.optimize_splines(m, dec, t, 1, M=2, seed=0, precision="tf32")
torch.cuda.synchronize()
lib = _lib.load()
out = (ctypes.c_longlong * (nc * 8))()
lib.vlg_debug_tc_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.vlg_debug_tc_stats(out, nc) == 0
st = np.array(out).reshape(nc, 8).astype(np.float64)
names = ["producer wait empty", "mma wait a_ready", "mma wait full", "mma total", "epi0 wait acc", "mma issue loops", "epi1 wait acc", "mma commits"]
for i, nm in enumerate(names):
    print(f"{nm:22s} mean {st[:, i].mean():12.0f} cycles  ({100 * st[:, i].mean() / st[:, 3].mean():5.1f}% of mma total)")

ph = (ctypes.c_longlong * (nc * 16))()
lib.vlg_debug_tc_phase.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.vlg_debug_tc_phase(ph, nc) == 0
ph = np.array(ph).reshape(nc, 16).astype(np.float64).mean(0)
items = 16 * 5  # fwd items of chain 0 per curve-step
names = ["sw wait+bar", "F1 compute+st", "F1 wait_st+arrive", "wait acc F2", "E-F2 tmem ld", "E-F2 compute+st", "E-F2 wait_st+arrive", "wait acc F3", "E-F3 ld+bias", "turn+Diff update"]
for i, nm in enumerate(names):
    print(f"fwd phase {nm:22s} {ph[i] / items:8.0f} cycles per item")
print("fwd item total", sum(ph[:10]) / items)
