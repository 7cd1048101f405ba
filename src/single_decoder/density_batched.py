"""python -m src.single_decoder.density_batched -- geodesic distance matrix of the single-decoder run (drop-in for
the matrix / JSON part of the reference's src/single_decoder/density_batched.py:41-142: same flags, same
geodesic_distances_seed*_p*.json).  Host-only bookkeeping; the density / heat-map figures of the reference are
plotting and are skipped when matplotlib is not installed."""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import torch

from vlg_b200 import formats


def distance_matrix_from_records(batch_data):
    """Points are numbered in order of first appearance (a then b of every entry), labels come from
    cluster_pair; matrix[a, b] = matrix[b, a] = length_geodesic, NaN where no spline exists, diagonal 0
    (density_batched.py:54-110)."""
    index_map, cluster_ids = {}, []
    for e in batch_data:
        for pt, label in zip((e["a"], e["b"]), e["cluster_pair"]):
            key = tuple(torch.as_tensor(pt).view(-1).tolist())
            if key not in index_map:
                index_map[key] = len(cluster_ids)
                cluster_ids.append(label)
    n = len(cluster_ids)
    mat = np.full((n, n), np.nan)
    for e in batch_data:
        ia = index_map[tuple(torch.as_tensor(e["a"]).view(-1).tolist())]
        ib = index_map[tuple(torch.as_tensor(e["b"]).view(-1).tolist())]
        mat[ia, ib] = mat[ib, ia] = e["length_geodesic"]
    np.fill_diagonal(mat, 0.0)
    return cluster_ids, mat


def main(seed, pairs_path, artifact_dir="src/artifacts"):
    pair_tag = Path(pairs_path).stem.replace("selected_pairs_", "")
    art = Path(artifact_dir)
    spline_path = art / f"spline_batch_optimized_batched_seed{seed}_p{pair_tag}.pt"
    json_path = art / f"geodesic_distances_seed{seed}_p{pair_tag}.json"
    batch_data = torch.load(spline_path, map_location="cpu", weights_only=False)
    print(f"Loaded {len(batch_data)} optimized splines from: {spline_path}")
    cluster_ids, mat = distance_matrix_from_records(batch_data)
    print(f"Distance matrix shape: {mat.shape}")
    formats.save_distance_json(seed, cluster_ids, mat, json_path)
    print(f"Saved: {json_path}")
    return json_path


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--seed", type=int, required=True, help="Seed used to identify model/data")
    parser.add_argument("--pairs_path", type=str, default="src/artifacts/selected_pairs.json", help="Path to selected pairs JSON")
    parser.add_argument("--artifact-dir", type=str, default="src/artifacts")
    args = parser.parse_args()
    main(args.seed, args.pairs_path, args.artifact_dir)
