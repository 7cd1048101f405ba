"""Drop-in for `python -m src.select_representative_pairs` (reference: src/select_representative_pairs.py): per
cell-type label, the data point whose latent code is nearest to the label's latent centroid; all pairs of these
representatives are the curves `src.init_splines_ensemble` / `src.optimize` then work on.  Same flags, same JSON
layout (`{"representatives": [{"index", "label"}], "pairs": [[i, j], ...]}`), same arithmetic (float32 centroid and
norms, first minimum wins) -- `tests/test_train_pairs.py` reproduces the reference's committed
`src/artifacts/selected_pairs_{10,50,133}.json` from its committed latents.  SURVEY §8 row f-4: host logic, one-off.
"""
from __future__ import annotations

import argparse
import json
from itertools import combinations
from pathlib import Path

import numpy as np
import torch


def select_representatives(latents, labels, max_labels: int = 10):
    """One representative per label for the first `max_labels` labels in sorted order
    (reference lines 22-38).  One stable sort groups the points; the groups keep index order, so the first minimum
    of a group is the reference's `argmin`."""
    latents = np.asarray(latents)
    labels = np.asarray(labels)
    names, first, counts = np.unique(labels, return_index=True, return_counts=True)
    if len(names) < max_labels:
        print(f"Warning: Only {len(names)} unique labels found, expected {max_labels}.")
    order = np.argsort(labels, kind="stable")
    starts = np.concatenate([[0], np.cumsum(counts)])
    reps = []
    for g in range(min(max_labels, len(names))):
        members = order[starts[g]:starts[g + 1]]
        pts = latents[members]
        dist = np.linalg.norm(pts - pts.mean(axis=0), axis=1)
        reps.append({"index": int(members[int(np.argmin(dist))]), "label": str(names[g])})
    return reps


def save_pairs(representatives, path):
    idx = [r["index"] for r in representatives]
    pairs = list(combinations(idx, 2))
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        json.dump({"representatives": representatives, "pairs": pairs}, f, indent=2)
    print(f"Saved {len(pairs)} pairs from {len(representatives)} representatives to {path}")


def load_pairs(path="src/artifacts/selected_pairs.json"):
    with open(path) as f:
        blob = json.load(f)
    return blob["representatives"], blob["pairs"]


def extract_latents(state_dict, data, device):
    """Encoder means of the whole data set (reference lines 16-20), through the functional encoder."""
    from vlg_b200 import evae
    with torch.no_grad():
        return evae.encoder_mean(state_dict, data.to(device)).cpu().numpy()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-type", choices=["vae", "evae"], required=True, help="Type of model: vae or evae")
    ap.add_argument("--latent-dim", type=int, default=2)
    ap.add_argument("--num-decoders", type=int, default=10, help="Only used for EVAE")
    ap.add_argument("--max-labels", type=int, default=10)
    ap.add_argument("--data-path", type=str, default="data/tasic-pca50.npy")
    ap.add_argument("--label-path", type=str, default="data/tasic-ttypes.npy")
    ap.add_argument("--vae-latent-path", type=str, default="src/artifacts/latents_VAE_ld2_ep100_bs64_lr1e-03_seed12.npy",
                    help="Only used for single VAE")
    ap.add_argument("--model-path", type=str, default="experiment/model_seed12.pt")
    ap.add_argument("--output-path", type=str, default="experiment/pairs/selected_pairs_10.json")
    args = ap.parse_args(argv)

    labels = np.load(args.label_path, allow_pickle=True)
    if args.model_type == "vae":
        print("[INFO] Using precomputed VAE latents from .npy file")
        latents = np.load(args.vae_latent_path)
    else:
        print("[INFO] Using EVAE model to extract latents via encoder")
        from vlg_b200 import evae
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        data = torch.from_numpy(np.load(args.data_path).astype(np.float32))
        latents = extract_latents(evae.load_state_dict(args.model_path), data, device)
    save_pairs(select_representatives(latents, labels, max_labels=args.max_labels), args.output_path)
    return args.output_path


if __name__ == "__main__":
    main()
