"""Cost of commits / accumulator switches / fresh accumulations in a tcgen05.mma stream."""
import sys, ctypes
sys.path.insert(0, ".")
import torch, vlg_b200
from vlg_b200 import _lib
lib = _lib.load()
fn = lib.vlg_selftest_mma_rate
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
ctas = 148
for mode, label in [(0, "plain stream"), (32, "commit every 4"), (64, "switch D every 8"), (128, "fresh accumulation every 16"), (224, "all three"),
                    (1, "plain + tmem ld/st traffic"), (225, "all three + tmem ld/st"), (227, "all three + tmem + lds")]:
    for iters in (16, 4096):
        out = torch.zeros(3 * ctas, dtype=torch.int64, device="cuda")
        assert fn(128, iters, 2048, 128, ctas, mode, out.data_ptr(), 0) == 0
        torch.cuda.synchronize()
        o = out.float().view(3, ctas).mean(1)
        print(f"{label:32s} iters {iters:5d}: total {o[0].item():9.0f} cycles = {o[0].item() / iters:6.1f} cycles/MMA")
