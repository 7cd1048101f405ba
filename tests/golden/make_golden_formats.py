"""Reference-produced FILES for the format tests: small committed outputs of the
reference are copied verbatim into tests/golden/ref_files/, and the distance matrices the reference's own
plot_geodesic_matrix (src/eval.py:13-66) builds from one of them are captured (seaborn.heatmap is stubbed
to record its argument) into ref_files/matrices.npz.

    python tests/golden/make_golden_formats.py        # build container only (/root/reference)
"""
import json
import shutil
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "ref_files"
COPIES = ["experiment/splines_opt_model_seed12/spline_batch_opt_euclidean_10.pt",
          "experiment/splines_init_model_seed12/spline_batch_init_euclidean_10.pt",
          "experiment/plots/cov_values_alldec_alldec.json",
          "experiment/pairs/selected_pairs_10.json",
          "src/artifacts/spline_batch_optimized_batched_seed12.pt"]

captured = {}
for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "seaborn", "mpl_toolkits",
             "mpl_toolkits.axes_grid1"]:
    sys.modules.setdefault(name, types.ModuleType(name))
plt = sys.modules["matplotlib.pyplot"]
for fn in ("figure", "xticks", "yticks", "title", "xlabel", "ylabel", "tight_layout", "savefig"):
    setattr(plt, fn, lambda *a, **k: None)
sys.modules["seaborn"].heatmap = lambda mat, **kw: captured.update(mat=np.array(mat), labels=list(kw["xticklabels"]))
sys.path.insert(0, str(REF))
import src.eval as ref_eval  # noqa: E402

if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    for rel in COPIES:
        shutil.copyfile(REF / rel, OUT / Path(rel).name)
    blob = torch.load(OUT / "spline_batch_opt_euclidean_10.pt", map_location="cpu", weights_only=False)
    mats = {}
    for len_type in ("geodesic", "euclidean_dist"):
        ref_eval.plot_geodesic_matrix(blob, "/tmp/x.png", len_type=len_type, seed=12, init_type="euclidean")
        mats[len_type] = captured["mat"]
        mats["labels"] = np.array([str(x) for x in captured["labels"]])
    # a blob with one pair missing and one foreign spline: NaN entry + "skipped" branch (src/eval.py:35-50)
    blob2 = {"spline_data": [dict(d) for d in blob["spline_data"][1:]], "representatives": blob["representatives"]}
    blob2["spline_data"][0]["a_index"] = -7
    ref_eval.plot_geodesic_matrix(blob2, "/tmp/x.png", len_type="geodesic", seed=12, init_type="euclidean")
    mats["geodesic_missing"] = captured["mat"]
    np.savez_compressed(OUT / "matrices.npz", **mats)
    # geodesic_distances_seed*_p*.json: the writer at src/single_decoder/density_batched.py:135-142 on a 12-cluster
    # slice of the committed matrix (the full 133 x 133 file is 449 KB)
    full = json.load(open(REF / "src/artifacts/geodesic_distances_seed12_p133.json"))
    n = 12
    json_matrix = {"seed": full["seed"], "cluster_ids": full["cluster_ids"][:n],
                   "distance_matrix": np.array(full["distance_matrix"])[:n, :n].tolist()}
    with open(OUT / "geodesic_distances_seed12_p12.json", "w") as f:
        json.dump(json_matrix, f, indent=2)
    print("wrote", sorted(p.name for p in OUT.iterdir()))
