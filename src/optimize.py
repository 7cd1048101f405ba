"""python -m src.optimize -- ensemble geodesic optimisation (drop-in for the reference's
src/optimize.py:80-244: same flags, same input/output files), on the B200 engine.

Differences a user can see: the whole pair list is optimised at once (--batch-size is accepted and
ignored: curves are independent, a batch only bounded memory in the reference); decoder draws come
from a counter-based generator keyed on --seed (the reference never seeds its draws); under
torchrun the pair list is sharded over the GPUs and rank 0 writes the file; plotting is skipped
when matplotlib is not installed; without data/tasic-pca50.npy the Euclidean distance falls back
to the latent end points stored in the spline file.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path

import numpy as np
import torch

import vlg_b200
from vlg_b200 import evae, formats, sharding


def main(model_path, spline_path, init_type, pair_count, steps=500, batch_size=200, M=2, precision=None, seed=0,
         data_path="data/tasic-pca50.npy", log_every=50):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("src.optimize needs a CUDA (B200) device: the engine has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    log = print if rank == 0 else (lambda *a, **k: None)
    log(f"[INFO] Device: {device} (world size {world})")

    if spline_path is None:
        model_name = Path(model_path).stem
        spline_path = Path("experiment") / f"splines_init_{model_name}" / f"spline_batch_init_{init_type}_{pair_count}.pt"
        if not spline_path.exists():
            raise FileNotFoundError(f"[ERROR] Expected spline file not found: {spline_path}")
        log(f"[INFO] Automatically using spline: {spline_path}")

    sd = evae.load_state_dict(model_path)
    decoders = vlg_b200.DecoderEnsemble.from_state_dict(sd, device)
    log(f"[DEBUG] Loaded model: {model_path} ({len(decoders)} decoders)")

    blob = formats.load_spline_blob(spline_path)
    spline_data = blob["spline_data"]
    arr = formats.splines_to_arrays(spline_data)
    n_poly, N = arr["n_poly"], len(spline_data)
    log(f"[INFO] Optimizing {N} splines (n_poly={n_poly}, K={arr['basis'].shape[1]})")
    t_vals = torch.linspace(0, 1, 2000, device=device)

    lo, hi = sharding.shard_range(N, rank, world)
    model = vlg_b200.GeodesicSplineBatch(arr["a"][lo:hi].to(device), arr["b"][lo:hi].to(device),
                                         arr["basis"].to(device), arr["omega"][lo:hi].to(device).clone(), n_poly)
    energy = torch.zeros(hi - lo, device=device)
    done = 0
    while done < steps:
        ns = min(log_every, steps - done)
        energy, trace = vlg_b200.optimize_splines(model, decoders, t_vals, ns, M=M, lr=1e-3, seed=seed, curve_id0=lo,
                                                  precision=precision, return_trace=True)
        if world == 1:
            log(f"[Step {done}] Mean Energy: {trace[0].mean().item():.4f}")
        done += ns
    omega_opt = sharding.gather_results(model.omega, N)
    lengths = sharding.gather_results(torch.sqrt(energy), N)  # src/optimize.py:168

    if rank == 0:
        if Path(data_path).exists():
            data = torch.tensor(np.load(data_path), dtype=torch.float32, device=device)
            ia = torch.tensor([d["a_index"] for d in spline_data], device=device)
            ib = torch.tensor([d["b_index"] for d in spline_data], device=device)
            with torch.no_grad():
                eucl = (evae.encoder_mean(sd, data[ia]) - evae.encoder_mean(sd, data[ib])).norm(dim=1).cpu()
        else:
            log(f"[WARNING] {data_path} not found: euclidean_distance uses the spline end points")
            eucl = (arr["a"] - arr["b"]).norm(dim=1)
        formats.write_back_optimized(spline_data, omega_opt, lengths, eucl)
        model_name = Path(model_path).stem
        spline_tag = Path(spline_path).stem.replace("spline_batch_init_", "")
        save_dir = Path("experiment") / f"splines_opt_{model_name}"
        save_path = save_dir / f"spline_batch_opt_{spline_tag}.pt"
        formats.save_opt_blob(spline_data, blob.get("representatives"), blob.get("pairs"), model_name, init_type,
                              pair_count, M, steps, save_path)
        print(f"[✓] Saved optimized splines to: {save_path}")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--model-path", type=str, required=True)
    parser.add_argument("--spline-path", type=str, default=None)
    parser.add_argument("--init-type", type=str, default="entropy", choices=["entropy", "euclidean"])
    parser.add_argument("--pair-count", type=int, required=True)
    parser.add_argument("--steps", type=int, default=100)
    parser.add_argument("--batch-size", type=int, default=200, help="accepted for compatibility; ignored")
    parser.add_argument("--mc-samples", type=int, default=2)
    parser.add_argument("--precision", type=str, default=vlg_b200.DEFAULT_PRECISION, choices=["f16", "f16x3", "f16x3f", "tf32", "fp32"],
                        help="tcgen05 tensor-core kernel: f16x3f (default: 3-term fp16 forward = fp32-grade energies, "
                             "single-term backward), f16x3 (all GEMMs 3-term: fp32-grade), f16 or tf32 operands "
                             "(<=1e-3 on lengths, fastest; fp16 operands must stay below 65504), or fp32: the CUDA-core kernel")
    parser.add_argument("--seed", type=int, default=0, help="seed of the decoder-pair draws")
    args = parser.parse_args()
    main(model_path=args.model_path, spline_path=args.spline_path, init_type=args.init_type, pair_count=args.pair_count,
         steps=args.steps, batch_size=args.batch_size, M=args.mc_samples, precision=args.precision, seed=args.seed)
