"""The six committed eVAE checkpoints of the CoV study (experiment/model_seed{12,123,1234,12345,45,456}.pt of the
reference) as one .npz of state-dict tensors (decoders as fp32, encoder as fp32), for the GPU tests of the CoV
driver: the reference checkout is not mounted on the GPU box.
data/tasic-pca50.npy is a missing blob, so the tests generate stand-in data points with the seed-12 ensemble
(mean of the ten decoders at random latent locations) and encode them with every seed's encoder, exactly
as src/eval.py:102-104 does with real cells.

    python tests/golden/make_golden_cov.py        # build container only (/root/reference)
"""
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
SEEDS = [12, 123, 1234, 12345, 45, 456]

if __name__ == "__main__":
    out = {"seeds": np.array(SEEDS)}
    for s in SEEDS:
        sd = torch.load(REF / f"experiment/model_seed{s}.pt", map_location="cpu", weights_only=True)
        for k, v in sd.items():
            if k.startswith("decoder.") or k.startswith("encoder."):
                out[f"s{s}/{k}"] = v.float().numpy()
    np.savez_compressed(OUT / "evae_six_seeds.npz", **out)
    print("wrote", OUT / "evae_six_seeds.npz", (OUT / "evae_six_seeds.npz").stat().st_size)
