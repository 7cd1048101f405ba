"""Checkpoint helpers for the reference's eVAE (src/train.py): the encoder mean (used for end
points / Euclidean distances) as plain torch functional calls on the checkpoint tensors.
The encoder is a one-off 50->256->128->4 MLP, not part of the hot path."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def load_state_dict(path, map_location="cpu"):
    return torch.load(path, map_location=map_location, weights_only=True)


def encoder_mean(state_dict, x: torch.Tensor) -> torch.Tensor:
    """``model.encoder(x).base_dist.loc`` (src/train.py:25-34,71-78): Linear-SiLU-LayerNorm x2,
    Linear, first half of the output is the mean."""
    g = lambda k: state_dict[f"encoder.encoder_net.{k}"].to(x.device, x.dtype)
    h = F.silu(F.linear(x, g("0.weight"), g("0.bias")))
    h = F.layer_norm(h, (h.shape[-1],), g("2.weight"), g("2.bias"))
    h = F.silu(F.linear(h, g("3.weight"), g("3.bias")))
    h = F.layer_norm(h, (h.shape[-1],), g("5.weight"), g("5.bias"))
    out = F.linear(h, g("6.weight"), g("6.bias"))
    return out[..., : out.shape[-1] // 2]


def num_decoders(state_dict) -> int:
    return 1 + max(int(k.split(".")[1]) for k in state_dict if k.startswith("decoder.") and k.split(".")[1].isdigit())
