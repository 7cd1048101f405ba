"""Shared helpers for the tests: golden loading, draw regeneration, decoder dicts."""
import zlib
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
DEC_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3")


def load(tag):
    return dict(np.load(GOLDEN / f"{tag}.npz"))


def decoder_arrays(g):
    """Decoder weight arrays of a golden case (its own, or the committed eVAE seed-12 ones)."""
    src = g if "W1" in g else load("evae_seed12_decoders")
    return {k: src[k] for k in DEC_KEYS}


def decoder_list(arrs, K, dtype):
    return [{k: arrs[k][i].astype(dtype) for k in DEC_KEYS} for i in range(K)]


def regen_draws(g):
    """Recorded draws are stored as (seed, shape, crc): torch's CPU generator is
    deterministic, so regenerate and verify the checksum."""
    shp = tuple(int(x) for x in g["draws_shape"])
    gen = torch.Generator().manual_seed(int(g["draw_seed"]))
    d = torch.randint(0, int(g["K"]), shp, generator=gen, dtype=torch.int64).numpy()
    assert zlib.crc32(d.astype(np.uint8).tobytes()) == int(g["draws_crc"]), "draw stream changed"
    return d


def tgrid(T, dtype=np.float32):
    return torch.linspace(0, 1, T).numpy().astype(dtype)


def relerr(x, ref):
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300))
