"""Golden for pair selection and eVAE training, made with the reference's OWN code:

* `select_representatives` (src/select_representative_pairs.py:22-38) on the reference's committed latents
  (`src/artifacts/latents_VAE_ld2_ep100_bs64_lr1e-03_seed12.npy`) and cell-type labels (`data/tasic-ttypes.npy`)
  for 10 / 50 / 133 labels -- these reproduce the committed `src/artifacts/selected_pairs_{10,50,133}.json`
  (checked here);
* 30 Adam steps of the reference's EVAE (src/train.py:48-65,91-104) on synthetic 50-d data, seed 7, 3 decoders:
  per-step loss and the final state dict.

    python tests/golden/make_golden_pairs.py        # build container only (/root/reference)
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "seaborn", "mpl_toolkits",
             "mpl_toolkits.axes_grid1"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, str(REF))
import src.select_representative_pairs as ref_sel  # noqa: E402
import src.train as ref_train  # noqa: E402


def synthetic_data(n=512, seed=7):
    g = torch.Generator().manual_seed(seed)
    centers = 20.0 * torch.randn(6, 50, generator=g)
    return (centers[torch.randint(0, 6, (n,), generator=g)] + 3.0 * torch.randn(n, 50, generator=g)).float()


def reference_training(data, seed, num_decoders, steps, batch, lr):
    torch.manual_seed(seed)
    np.random.seed(seed)
    enc = ref_train.GaussianEncoder(ref_train.make_encoder_net(50, 2))
    dec = ref_train.GaussianDecoder(ref_train.make_decoder_net(2, 50))
    model = ref_train.EVAE(ref_train.GaussianPrior(2), enc, dec, num_decoders=num_decoders, beta=1.0)
    init = {k: v.clone() for k, v in model.state_dict().items()}
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for s in range(steps):
        x = data[(s * batch) % len(data):(s * batch) % len(data) + batch]
        loss = model(x)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return init, model.state_dict(), np.array(losses)


if __name__ == "__main__":
    lat = np.load(REF / "src/artifacts/latents_VAE_ld2_ep100_bs64_lr1e-03_seed12.npy")
    lab = np.load(REF / "data/tasic-ttypes.npy", allow_pickle=True)
    out = {"latents": lat, "labels": lab.astype("U")}
    for n in (10, 50, 133):
        reps = ref_sel.select_representatives(lat, lab, max_labels=n)
        committed = json.load(open(REF / f"src/artifacts/selected_pairs_{n}.json"))
        assert reps == committed["representatives"], n
        assert [list(p) for p in __import__("itertools").combinations([r["index"] for r in reps], 2)] == committed["pairs"]
        out[f"rep_index_{n}"] = np.array([r["index"] for r in reps])
        out[f"rep_label_{n}"] = np.array([r["label"] for r in reps])
    np.savez_compressed(OUT / "pairs_seed12.npz", **out)
    (OUT / "ref_files" / "artifacts_selected_pairs_10.json").write_bytes((REF / "src/artifacts/selected_pairs_10.json").read_bytes())

    data = synthetic_data()
    init, final, losses = reference_training(data, seed=7, num_decoders=3, steps=30, batch=64, lr=1e-3)
    blob = {"data": data.numpy(), "losses": losses}
    blob.update({"init/" + k: v.numpy() for k, v in init.items()})
    blob.update({"final/" + k: v.numpy() for k, v in final.items()})
    np.savez_compressed(OUT / "evae_train_30.npz", **blob)
    print("pairs_seed12.npz", (OUT / "pairs_seed12.npz").stat().st_size, "evae_train_30.npz", (OUT / "evae_train_30.npz").stat().st_size,
          "losses", losses[:3], losses[-1])
