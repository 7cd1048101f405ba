"""TMEM ld/st round-trip latency with and without a concurrent tcgen05.mma stream."""
import sys, ctypes
sys.path.insert(0, ".")
import torch, vlg_b200
from vlg_b200 import _lib
vlg_b200.build.build_selftest(); lib = _lib.load_selftest()
fn = lib.vlg_selftest_mma_rate
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
ctas = 148
for mode, iters, label in [(4, 0, "8 warps, no MMA"), (12, 0, "4 warps, no MMA"), (4, 4096, "8 warps, MMA stream"), (12, 4096, "4 warps, MMA stream"), (28, 64, "single MMA latency")]:
    out = torch.zeros(3 * ctas, dtype=torch.int64, device="cuda")
    assert fn(128, iters, 2048, 128, ctas, mode, out.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    o = out.float().view(3, ctas).mean(1)
    print(f"{label:22s}: ld+st round trip {o[1].item() / 256:7.1f} cycles; MMA stream {o[0].item() / max(iters, 1):6.1f} cycles/MMA; single MMA+commit {o[2].item():.0f}")
