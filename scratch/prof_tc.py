"""Small profiling driver: one launch of the step kernel on NC curves (one wave)."""
import sys
sys.path.insert(0, ".")
import torch, numpy as np
import vlg_b200, bench
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 148
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
w, a, b, om, _ = bench.synthetic_workload(nc)
dev = "cuda"
dec = vlg_b200.DecoderEnsemble.from_arrays(*[w[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")], dev)
basis, _ = vlg_b200.construct_nullspace_basis(4)
m = vlg_b200.GeodesicSplineBatch(a.to(dev), b.to(dev), basis.to(dev), om.to(dev), 4)
t = torch.linspace(0, 1, 2000, device=dev)
for _ in range(2):
    e = vlg_b200.optimize_splines(m, dec, t, steps, M=2, seed=0, precision=prec)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); e = vlg_b200.optimize_splines(m, dec, t, steps, M=2, seed=0, precision=prec); ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
print(f"{prec}: {nc} curves x {steps} steps: {ms:.3f} ms -> {nc*steps/ms*1e3:.0f} spline-steps/s, E[0]={float(e[0]):.2f}")
