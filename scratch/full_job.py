"""The north-star job end to end on one B200: 8778 pairs, 10 decoders, T=2000, M=2, 1000 Adam steps.
Times the tensor-core modes and compares their final lengths with the fp32 kernel's, curve by curve."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch, vlg_b200, bench
dev = "cuda"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w, a, b, omega, weights = bench.synthetic_workload(bench.N_CURVES)
dec = vlg_b200.DecoderEnsemble.from_arrays(*[w[k] for k in ("W1", "b1", "W2", "b2", "W3", "b3")], dev)
basis, _ = vlg_b200.construct_nullspace_basis(4)
t = torch.linspace(0, 1, 2000, device=dev)
out = {}
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f16", "tf32", "fp32"]
for prec in modes:
    m = vlg_b200.GeodesicSplineBatch(a.to(dev), b.to(dev), basis.to(dev), omega.to(dev), 4)
    vlg_b200.optimize_splines(vlg_b200.GeodesicSplineBatch(a.to(dev), b.to(dev), basis.to(dev), omega.to(dev), 4), dec, t, 1, M=2, seed=0, precision=prec)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    done = 0
    while done < steps:   # the drop-in CLI launches in chunks of 100 steps too
        ns = min(100, steps - done)
        e = vlg_b200.optimize_splines(m, dec, t, ns, M=2, seed=0, precision=prec)
        done += ns
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[prec] = (np.sqrt(e.cpu().numpy().astype(np.float64)), m.omega.cpu().numpy(), dt)
    print(f"{prec:5s}: {bench.N_CURVES} curves x {steps} steps in {dt:7.2f} s -> {bench.N_CURVES * steps / dt:9.0f} spline-steps/s; mean length {out[prec][0].mean():.4f}", flush=True)
ref = out["fp32"][0]
res = {"steps": steps, "weights": weights}
for prec in [m_ for m_ in modes if m_ != "fp32"]:
    rel = np.abs(out[prec][0] / ref - 1)
    res[prec] = {"seconds": out[prec][2], "max_rel_length_diff_vs_fp32": float(rel.max()), "mean_rel_length_diff_vs_fp32": float(rel.mean()),
                 "p999": float(np.quantile(rel, 0.999))}
    print(f"{prec}: final length vs fp32 kernel over {len(ref)} curves: max {rel.max():.2e}, 99.9 % {np.quantile(rel, 0.999):.2e}, mean {rel.mean():.2e}")
res["fp32"] = {"seconds": out["fp32"][2]}
open("gpurun_out/full_job.json", "w").write(json.dumps(res, indent=1))
np.savez_compressed("gpurun_out/full_job_lengths.npz", **{f"len_{k}": v[0].astype(np.float32) for k, v in out.items()},
                    **{f"omega_{k}": v[1] for k, v in out.items()})
