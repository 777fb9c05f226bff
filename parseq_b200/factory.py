"""Model factory with the reference's entry points: `create_model`, `load_from_checkpoint`,
`parse_model_args` (strhub/models/utils.py:73-104) for the PARSeq experiments
(configs/experiment/parseq*.yaml) and ViTSTR (configs/experiment/vitstr.yaml), which shares the ViT encoder kernels.
The CNN / RNN model families of the reference (ABINet, CRNN, TRBA) are out of scope."""
from __future__ import annotations

from typing import Any

import torch

from .config import PRESETS, make_config
from .system import InvalidModelError, PARSeq, ViTSTR

_WEIGHTS_URL = {   # strhub/models/utils.py:14-22 (PARSeq entries)
    "parseq-tiny": "https://github.com/baudm/parseq/releases/download/v1.0.0/parseq_tiny-e7a21b54.pt",
    "parseq-patch16-224": "https://github.com/baudm/parseq/releases/download/v1.0.0/parseq_small_patch16_224-fcf06f5a.pt",
    "parseq": "https://github.com/baudm/parseq/releases/download/v1.0.0/parseq-bb5792a6.pt",
    "vitstr": "https://github.com/baudm/parseq/releases/download/v1.0.0/vitstr-26d0fcf4.pt",
}


def _get_model_class(key: str):
    """strhub/models/utils.py:47-62 for the families served here."""
    if "parseq" in key:
        return PARSeq
    if "vitstr" in key:
        return ViTSTR
    raise InvalidModelError(f"Unable to find model class for '{key}'")


def get_pretrained_weights(experiment: str):
    if experiment not in _WEIGHTS_URL:
        raise InvalidModelError(f"No pretrained weights found for '{experiment}'")
    return torch.hub.load_state_dict_from_url(url=_WEIGHTS_URL[experiment], map_location="cpu", check_hash=True)


def create_model(experiment: str, pretrained: bool = False, **kwargs: Any):
    if experiment not in PRESETS:
        _get_model_class(experiment)
        raise InvalidModelError(f"No configuration found for '{experiment}'")
    cls = _get_model_class(experiment)
    cfg = make_config(experiment, **kwargs)
    kw = cfg.to_kwargs()
    kw.update(cfg.extra)
    if cls is ViTSTR:        # ctor kwargs of vitstr/system.py:31-46 (everything else lands in **kwargs there too)
        kw["num_heads"] = kw.pop("enc_num_heads")
        for k in ("arch", "enc_mlp_ratio", "enc_depth", "dec_num_heads", "dec_mlp_ratio", "dec_depth", "decode_ar",
                  "refine_iters", "perm_num", "perm_forward", "perm_mirrored", "dropout"):
            kw.pop(k, None)
    else:
        kw.pop("arch", None)
    model = cls(**kw)
    if pretrained:
        # strhub/models/utils.py:80-82: PARSeq weights belong to the inner model, the others to the system
        m = model.model if cls is PARSeq else model
        m.load_state_dict(get_pretrained_weights(experiment))
    return model


def load_from_checkpoint(checkpoint_path: str, **kwargs: Any) -> PARSeq:
    if checkpoint_path.startswith("pretrained="):
        return create_model(checkpoint_path.split("=", maxsplit=1)[1], True, **kwargs)
    return _get_model_class(checkpoint_path).load_from_checkpoint(checkpoint_path, **kwargs)


def parse_model_args(args):
    """`name:type=value` CLI overrides (README.md:180)."""
    casts = {"int": int, "float": float, "str": str, "bool": lambda v: v.lower() == "true"}
    out = {}
    for arg in args:
        lhs, value = arg.split("=", maxsplit=1)
        name, tname = lhs.split(":", maxsplit=1)
        out[name] = casts[tname](value)
    return out
