"""Issue-to-completion cost of a tcgen05.mma stream for M = 128 / 64, kind::tf32 / kind::f16, N = 64 / 128."""
import sys, ctypes
sys.path.insert(0, ".")
import torch, vlg_b200
from vlg_b200 import _lib
lib = _lib.load()
fn = lib.vlg_selftest_mma_rate
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
ctas = 148
for kind, kb in (("tf32", 0), ("f16", 512)):
    for M, mb in ((128, 0), (64, 256)):
        for (N, lbo) in ((128, 2048), (64, 1024)):
            out = torch.zeros(3 * ctas, dtype=torch.int64, device="cuda")
            iters = 4096
            assert fn(N, iters, lbo, 128, ctas, kb | mb, out.data_ptr(), 0) == 0
            torch.cuda.synchronize()
            print(f"kind::{kind:4s} M={M:3d} N={N:3d}: {out.float().view(3, ctas).mean(1)[0].item() / iters:6.1f} cycles/MMA")
