"""On-disk layouts of the reference, byte-compatible readers / writers.

  spline_batch_init_*.pt   {"spline_data": [dict...], "representatives": [...], "pairs": [...]}
                           (src/init_splines_ensemble.py:195-216)
  spline_batch_opt_*.pt    same + per-spline omega_optimized / geodesic_length / euclidean_distance
                           and "metadata" (src/optimize.py:182-201)
  single-decoder list      [dict(a,b,cluster_pair,n_poly,basis,omega_init,omega_optimized,
                           length_geodesic,length_euclidean)] (optimize_energy_batched.py:108-124)
  geodesic_distances JSON  {"seed","cluster_ids","distance_matrix"} (density_batched.py:135-142)
  cov_values JSON          (src/eval.py:145-157)
  pairs JSON               {"representatives":[{"index","label"}], "pairs":[[i,j]...]}
                           (src/select_representative_pairs.py:37-49)

The kernels want structure-of-arrays; the files are lists of dicts.  Pure host code.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch


def load_pairs(path) -> tuple:
    """(representatives, pairs) -- same return as the reference's load_pairs."""
    with open(path, "r") as f:
        data = json.load(f)
    return data["representatives"], data["pairs"]


def save_pairs(representatives: List[dict], pairs, path) -> None:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        json.dump({"representatives": representatives, "pairs": [list(p) for p in pairs]}, f, indent=2)


def load_spline_blob(path, map_location="cpu") -> dict:
    blob = torch.load(path, map_location=map_location, weights_only=False)
    if isinstance(blob, list):  # single-decoder scripts store a bare list
        blob = {"spline_data": blob, "representatives": None, "pairs": None}
    return blob


def splines_to_arrays(spline_data: List[dict], omega_key: str = "omega_init") -> Dict[str, torch.Tensor]:
    """list-of-dicts -> SoA: a [N,2], b [N,2], omega [N,Kb,2], basis [4n,Kb], n_poly.
    The basis is taken from the file (never recomputed, SURVEY hard part 7)."""
    return {
        "a": torch.stack([torch.as_tensor(d["a"]) for d in spline_data]).float(),
        "b": torch.stack([torch.as_tensor(d["b"]) for d in spline_data]).float(),
        "omega": torch.stack([torch.as_tensor(d[omega_key]) for d in spline_data]).float(),
        "basis": torch.as_tensor(spline_data[0]["basis"]).float(),
        "n_poly": int(spline_data[0]["n_poly"]),
    }


def init_spline_dict(a, b, a_index, b_index, a_label, b_label, n_poly, basis, omega_init) -> dict:
    """One entry of spline_batch_init_*.pt (src/init_splines_ensemble.py:195-205), CPU tensors."""
    return {"a": torch.as_tensor(a).detach().cpu(), "b": torch.as_tensor(b).detach().cpu(), "a_index": a_index,
            "b_index": b_index, "a_label": a_label, "b_label": b_label, "n_poly": int(n_poly),
            "basis": torch.as_tensor(basis).detach().cpu(), "omega_init": torch.as_tensor(omega_init).detach().cpu()}


def save_init_blob(spline_data, representatives, pairs, path) -> None:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    torch.save({"spline_data": spline_data, "representatives": representatives, "pairs": pairs}, path)


def write_back_optimized(spline_data: List[dict], omega_optimized: torch.Tensor, geodesic_length: torch.Tensor,
                         euclidean_distance) -> None:
    """Per-spline results into the dicts, as src/optimize.py:182-185 does."""
    om = omega_optimized.detach().cpu()
    gl = geodesic_length.detach().cpu()
    for i, d in enumerate(spline_data):
        d["omega_optimized"] = om[i]
        d["geodesic_length"] = float(gl[i])
        d["euclidean_distance"] = float(euclidean_distance[i])


def save_opt_blob(spline_data, representatives, pairs, model_name, init_type, pair_count, mc_samples, steps,
                  path) -> None:
    """spline_batch_opt_*.pt (src/optimize.py:190-201)."""
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    torch.save({"spline_data": spline_data, "representatives": representatives, "pairs": pairs,
                "metadata": {"model_name": model_name, "init_type": init_type, "pair_count": pair_count,
                             "mc_samples": mc_samples, "steps": steps}}, path)


def single_decoder_records(a, b, cluster_pairs, n_poly, basis, omega_init, omega_optimized, length_geodesic) -> list:
    """The list written by src/single_decoder/optimize_energy_batched.py:108-124."""
    out = []
    for i in range(a.shape[0]):
        out.append({"a": a[i].cpu(), "b": b[i].cpu(), "cluster_pair": cluster_pairs[i], "n_poly": int(n_poly),
                    "basis": basis.cpu(), "omega_init": omega_init[i].cpu(), "omega_optimized": omega_optimized[i].cpu(),
                    "length_geodesic": float(length_geodesic[i]), "length_euclidean": float(torch.norm(a[i] - b[i]))})
    return out


def distance_matrix(spline_blob: dict, len_type: str = "geodesic") -> tuple:
    """Symmetric NaN-filled matrix over the representatives, diagonal 0
    (src/eval.py:13-50).  Returns (matrix [n,n], labels, skipped)."""
    reps = spline_blob["representatives"]
    if reps is None:
        raise ValueError("Missing 'representatives' in spline blob. Cannot build label mapping.")
    g2l = {r["index"]: i for i, r in enumerate(reps)}
    labels = [r.get("cluster_label") or r.get("label") or r.get("index") or str(i) for i, r in enumerate(reps)]
    n = len(reps)
    mat = np.full((n, n), np.nan)
    skipped = 0
    key = "geodesic_length" if len_type == "geodesic" else "euclidean_distance"
    for d in spline_blob["spline_data"]:
        ia, ib = d["a_index"], d["b_index"]
        if ia not in g2l or ib not in g2l:
            skipped += 1
            continue
        mat[g2l[ia], g2l[ib]] = mat[g2l[ib], g2l[ia]] = d[key]
    np.fill_diagonal(mat, 0)
    return mat, labels, skipped


def save_distance_json(seed: int, cluster_ids, matrix: np.ndarray, path) -> None:
    """geodesic_distances_seed*_p*.json (src/single_decoder/density_batched.py:135-142)."""
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        json.dump({"seed": seed, "cluster_ids": list(cluster_ids), "distance_matrix": np.asarray(matrix).tolist()}, f, indent=2)


def cov_payload(avg_cov_geo: dict, avg_cov_euc: float, raw_geo: dict, raw_euc: list, seeds, decoder_counts,
                num_pairs: int) -> dict:
    """cov_values_*.json content (src/eval.py:145-153)."""
    return {"avg_cov_geodesic": {str(k): float(v) for k, v in avg_cov_geo.items()},
            "avg_cov_euclidean": float(avg_cov_euc),
            "raw_cov_geodesic": {str(k): [float(x) for x in v] for k, v in raw_geo.items()},
            "raw_cov_euclidean": [float(x) for x in raw_euc], "seeds": list(seeds),
            "decoder_counts": list(decoder_counts), "num_pairs": int(num_pairs)}
