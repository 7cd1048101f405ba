"""python -m src.eval -- distance matrix / CoV study (drop-in for the reference's src/eval.py:
same flags and output files).  `--mode cov` batches all pairs of ALL seeds of a decoder count k into one
launch of the engine (the reference runs 6300 sequential B=1 optimisations, src/eval.py:90-128)."""
from __future__ import annotations

import argparse
import json
from pathlib import Path

import numpy as np
import torch

import vlg_b200
from vlg_b200 import evae, formats


def compute_cov(values):
    values = np.array(values)
    return np.std(values) / np.mean(values) if np.mean(values) > 0 else 0.0


def plot_geodesic_matrix(spline_blob, output_path, len_type="geodesic", seed=None, init_type=None):
    mat, labels, skipped = formats.distance_matrix(spline_blob, len_type)
    if skipped:
        print(f"[INFO] Skipped {skipped} spline entries not in representative set")
    np.save(Path(output_path).with_suffix(".npy"), mat)
    try:
        import matplotlib.pyplot as plt
        import seaborn as sns
    except ImportError:
        print(f"[INFO] matplotlib/seaborn not installed: wrote {Path(output_path).with_suffix('.npy')} only")
        return mat
    plt.figure(figsize=(10, 10))
    sns.heatmap(mat, square=True, xticklabels=labels, yticklabels=labels, cmap="copper", cbar=False)
    plt.title(f"Geodesic Distance Matrix - seed {seed} (init by {init_type})" if len_type == "geodesic"
              else f"Euclidean Distance Matrix - seed {seed}")
    plt.tight_layout()
    plt.savefig(output_path, dpi=300)
    print(f"[✓] Saved geodesic matrix plot to: {output_path}")
    return mat


def cov_lengths(state_dicts, za, zb, decoder_counts, steps=300, precision=None, draw_seed=0, device="cuda"):
    """sqrt(energy) after `steps` Adam steps from omega = 0 for every (seed, pair, k) of the CoV study
    (src/eval.py:108-128), ONE launch per k: the ensembles of all seeds are packed into one buffer and every
    curve carries the index of its own weight set (decoder_base), so 6 seeds x 105 pairs x 10 decoder counts are
    10 launches instead of the reference's 6300 sequential B=1 optimisations.
    za, zb: [S, N, 2] end points per seed.  Returns {k: array [S, N]}."""
    S, N = za.shape[0], za.shape[1]
    per_set = evae.num_decoders(state_dicts[0])
    decoders = vlg_b200.DecoderEnsemble.from_state_dicts(state_dicts, device, num_decoders=per_set)
    basis, _ = vlg_b200.construct_nullspace_basis(4, device)
    t_vals = torch.linspace(0, 1, 2000, device=device)
    a = za.reshape(S * N, 2).to(device).float().contiguous()
    b = zb.reshape(S * N, 2).to(device).float().contiguous()
    base = (torch.arange(S, dtype=torch.int32) * per_set).repeat_interleave(N)
    geo = {}
    for k in decoder_counts:
        model = vlg_b200.GeodesicSplineBatch(a, b, basis, torch.zeros((S * N, basis.shape[1], 2), device=device), 4)
        # a single decoder is the deterministic energy: 11-bit operands cannot resolve it (the engine then runs the
        # fp32 kernel; the 3-term mode can stay on the tensor pipe)
        energy = vlg_b200.optimize_splines(model, decoders, t_vals, steps, M=2, seed=draw_seed, precision=precision,
                                           decoder_base=base, k_active=k)
        geo[k] = torch.sqrt(energy).view(S, N).cpu().numpy()
    return geo


def run_cov_analysis(seeds, decoder_counts, pairfile, model_dir, data_path, output_plot, steps=300, precision=None,
                     draw_seed=0):
    """CoV of geodesic lengths across seeds for k = 1..10 decoders (src/eval.py:74-159)."""
    device = torch.device("cuda")
    data = torch.tensor(np.load(data_path).astype(np.float32), device=device)
    _, pairs = formats.load_pairs(pairfile)
    ia = torch.tensor([p[0] for p in pairs], device=device)
    ib = torch.tensor([p[1] for p in pairs], device=device)
    N = len(pairs)
    sds = [evae.load_state_dict(f"{model_dir}/model_seed{seed}.pt") for seed in seeds]
    with torch.no_grad():   # end points = encoder means of the two cells under each seed's encoder (src/eval.py:102-104)
        za = torch.stack([evae.encoder_mean(sd, data[ia]) for sd in sds])
        zb = torch.stack([evae.encoder_mean(sd, data[ib]) for sd in sds])
    euc = (za - zb).norm(dim=2).cpu().numpy()
    geo = cov_lengths(sds, za, zb, decoder_counts, steps, precision, draw_seed, device)
    cov_geo = {k: [compute_cov(geo[k][:, i]) for i in range(N)] for k in decoder_counts}
    cov_euc = [compute_cov(euc[:, i]) for i in range(N)]
    payload = formats.cov_payload({k: np.mean(cov_geo[k]) for k in decoder_counts}, np.mean(cov_euc), cov_geo, cov_euc,
                                  seeds, decoder_counts, N)
    json_path = Path(output_plot).with_name(f"cov_values_alldec_{Path(output_plot).stem.split('_')[-1]}.json")
    json_path.parent.mkdir(parents=True, exist_ok=True)
    with open(json_path, "w") as f:
        json.dump(payload, f, indent=2)
    print(f"[✓] Saved CoV values to: {json_path}")
    return payload


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--mode", type=str, choices=["matrix", "cov"], required=True)
    parser.add_argument("--len-type", type=str, default="geodesic", choices=["geodesic", "euclidean_dist"])
    parser.add_argument("--init-type", type=str, default=None, choices=["entropy", "euclidean"])
    parser.add_argument("--pair-count", type=int, default=133)
    parser.add_argument("--seed", type=int)
    parser.add_argument("--seeds", nargs="*", type=int, default=[12, 123])
    parser.add_argument("--precision", type=str, default=vlg_b200.DEFAULT_PRECISION, choices=["f16", "f16x3", "f16x3f", "tf32", "fp32"])
    args = parser.parse_args()
    plot_dir = Path("experiment/plots")
    plot_dir.mkdir(parents=True, exist_ok=True)
    if args.mode == "matrix":
        spline_path = Path("experiment") / f"splines_opt_model_seed{args.seed}" / f"spline_batch_opt_{args.init_type}_{args.pair_count}.pt"
        if not spline_path.exists():
            print(f"[ERROR] File not found: {spline_path}")
            return
        blob = formats.load_spline_blob(spline_path)
        if args.len_type == "geodesic" and args.init_type is None:
            raise ValueError("For geodesic length, --init-type must be specified.")
        plot_path = plot_dir / f"{args.len_type}_matrix_seed{args.seed}_{args.init_type or ''}_{args.pair_count}.png"
        plot_geodesic_matrix(blob, plot_path, len_type=args.len_type, seed=args.seed, init_type=args.init_type)
    else:
        run_cov_analysis(seeds=args.seeds, decoder_counts=list(range(1, 11)),
                         pairfile=f"experiment/pairs/selected_pairs_{args.pair_count}.json", model_dir="experiment",
                         data_path="data/tasic-pca50.npy", output_plot=f"experiment/plots/cov_plot_{args.pair_count}_alldec.png",
                         precision=args.precision)


if __name__ == "__main__":
    main()
