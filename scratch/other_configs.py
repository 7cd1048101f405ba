"""Throughput of the other BASELINE configs (2: single decoder, all 8778 splines; 5: 64 decoders, n_poly 8, T=256)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch, vlg_b200, bench
dev = "cuda"
def run(label, N, K, T, n_poly, M, steps, prec, w=None):
    rng = np.random.default_rng(0)
    if w is None:
        w = dict(W1=rng.normal(size=(K, 128, 2)) * 0.7, b1=rng.normal(size=(K, 128)) * 0.3, W2=rng.normal(size=(K, 128, 128)) * 0.09,
                 b2=rng.normal(size=(K, 128)) * 0.1, W3=rng.normal(size=(K, 50, 128)) * 0.09, b3=rng.normal(size=(K, 50)) * 0.1)
        w = {k: torch.tensor(v, dtype=torch.float32) for k, v in w.items()}
    dec = vlg_b200.DecoderEnsemble.from_arrays(*[w[k][:K] for k in ("W1", "b1", "W2", "b2", "W3", "b3")], dev)
    basis, _ = vlg_b200.construct_nullspace_basis(n_poly)
    a = torch.tensor(rng.uniform(-3, 3, (N, 2)), dtype=torch.float32); b = torch.tensor(rng.uniform(-3, 3, (N, 2)), dtype=torch.float32)
    om = torch.tensor(0.1 * rng.normal(size=(N, n_poly + 1, 2)), dtype=torch.float32)
    m = vlg_b200.GeodesicSplineBatch(a.to(dev), b.to(dev), basis.to(dev), om.to(dev), n_poly)
    t = torch.linspace(0, 1, T, device=dev)
    for _ in range(2): vlg_b200.optimize_splines(m, dec, t, 1, M=M, seed=0, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); vlg_b200.optimize_splines(m, dec, t, steps, M=M, seed=0, precision=prec); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{label:46s} {prec:5s}: {N} curves x {steps} steps in {ms:8.2f} ms -> {N * steps / ms * 1e3:10.0f} spline-steps/s")
w3, *_ = bench.synthetic_workload(8)
run("config 2: single decoder, 8778 splines, T=2000", 8778, 1, 2000, 4, 1, 4, "fp32", w3)
for prec in ("f16", "tf32", "fp32"):
    run("config 5: 64 decoders, n_poly 8, T=256", 20000, 64, 256, 8, 2, 4, prec)
