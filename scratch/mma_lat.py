import sys, ctypes
sys.path.insert(0, ".")
import torch, vlg_b200
from vlg_b200 import _lib
lib = _lib.load()
fn = lib.vlg_selftest_mma_rate
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
for iters in (1, 2, 4, 8, 16, 32, 64):
    out = torch.zeros(148, dtype=torch.int64, device="cuda")
    for rep in range(3):
        assert fn(128, iters, 2048, 128, 148, 0, out.data_ptr(), 0) == 0
        torch.cuda.synchronize()
    print(f"iters {iters:3d}: {out.float().mean().item():.0f} cycles from first issue to mbarrier-observed completion")
