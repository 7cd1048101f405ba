"""Why is the e2e leg slower than the resident leg?  Times the pieces separately."""
import sys, time
sys.path.insert(0, ".")
import torch, vlg_b200, bench
prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
dev = torch.device("cuda", 0)
w, a, b, omega, _ = bench.synthetic_workload(bench.N_CURVES)
dec = vlg_b200.DecoderEnsemble.from_arrays(*[w[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")], dev)
basis, _ = vlg_b200.construct_nullspace_basis(4); basis = basis.to(dev)
t = torch.linspace(0, 1, 2000, device=dev)
h_a, h_b, h_om = (x.contiguous().pin_memory() for x in (a, b, omega))
def timed(fn, label):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{label:50s} {e0.elapsed_time(e1):8.2f} ms (wall {1e3*(time.perf_counter()-t0):8.2f})")
model = vlg_b200.GeodesicSplineBatch(h_a.to(dev), h_b.to(dev), basis, h_om.to(dev), 4)
for _ in range(3): vlg_b200.optimize_splines(model, dec, t, 1, M=2, seed=0, precision=prec)
timed(lambda: vlg_b200.optimize_splines(model, dec, t, 10, M=2, seed=0, precision=prec), "resident, 10 steps (after 3 warm steps)")
timed(lambda: vlg_b200.optimize_splines(model, dec, t, 10, M=2, seed=0, precision=prec), "resident, 10 steps again")
def fresh():
    m2 = vlg_b200.GeodesicSplineBatch(h_a.to(dev, non_blocking=True), h_b.to(dev, non_blocking=True), basis, h_om.to(dev, non_blocking=True), 4)
    return vlg_b200.optimize_splines(m2, dec, t, 10, M=2, seed=0, precision=prec)
timed(fresh, "fresh state from host, 10 steps")
timed(fresh, "fresh state from host, 10 steps again")
m3 = vlg_b200.GeodesicSplineBatch(h_a.to(dev), h_b.to(dev), basis, h_om.to(dev), 4)
timed(lambda: vlg_b200.optimize_splines(m3, dec, t, 10, M=2, seed=0, precision=prec), "fresh state resident, 10 steps")
timed(lambda: vlg_b200.optimize_splines(m3, dec, t, 10, M=2, seed=0, precision=prec), "same model, next 10 steps")
