"""Model factory with the reference's entry points: `create_model`, `load_from_checkpoint`,
`parse_model_args` (strhub/models/utils.py:73-104) for the PARSeq experiments
(configs/experiment/parseq*.yaml).  Other model families of the reference are out of scope."""
from __future__ import annotations

from typing import Any

import torch

from .config import PRESETS, make_config
from .system import InvalidModelError, PARSeq

_WEIGHTS_URL = {   # strhub/models/utils.py:14-22 (PARSeq entries)
    "parseq-tiny": "https://github.com/baudm/parseq/releases/download/v1.0.0/parseq_tiny-e7a21b54.pt",
    "parseq-patch16-224": "https://github.com/baudm/parseq/releases/download/v1.0.0/parseq_small_patch16_224-fcf06f5a.pt",
    "parseq": "https://github.com/baudm/parseq/releases/download/v1.0.0/parseq-bb5792a6.pt",
}


def get_pretrained_weights(experiment: str):
    if experiment not in _WEIGHTS_URL:
        raise InvalidModelError(f"No pretrained weights found for '{experiment}'")
    return torch.hub.load_state_dict_from_url(url=_WEIGHTS_URL[experiment], map_location="cpu", check_hash=True)


def create_model(experiment: str, pretrained: bool = False, **kwargs: Any) -> PARSeq:
    if experiment not in PRESETS:
        if "parseq" not in experiment:
            raise InvalidModelError(f"Unable to find model class for '{experiment}'")
        raise InvalidModelError(f"No configuration found for '{experiment}'")
    cfg = make_config(experiment, **kwargs)
    kw = cfg.to_kwargs()
    kw.update(cfg.extra)
    model = PARSeq(**kw)
    if pretrained:
        model.model.load_state_dict(get_pretrained_weights(experiment))
    return model


def load_from_checkpoint(checkpoint_path: str, **kwargs: Any) -> PARSeq:
    if checkpoint_path.startswith("pretrained="):
        return create_model(checkpoint_path.split("=", maxsplit=1)[1], True, **kwargs)
    if "parseq" not in checkpoint_path:
        raise InvalidModelError(f"Unable to find model class for '{checkpoint_path}'")
    return PARSeq.load_from_checkpoint(checkpoint_path, **kwargs)


def parse_model_args(args):
    """`name:type=value` CLI overrides (README.md:180)."""
    casts = {"int": int, "float": float, "str": str, "bool": lambda v: v.lower() == "true"}
    out = {}
    for arg in args:
        lhs, value = arg.split("=", maxsplit=1)
        name, tname = lhs.split(":", maxsplit=1)
        out[name] = casts[tname](value)
    return out
