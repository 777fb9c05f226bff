"""Architecture / decode configuration of the PARSeq inference path.

Field names follow the ctor kwargs of the reference system class
(/root/reference/strhub/models/parseq/system.py:35-60) and the hydra files
configs/model/parseq.yaml:5-25, configs/main.yaml:9-17,
configs/experiment/parseq-tiny.yaml:5-9, configs/experiment/parseq-patch16-224.yaml:5-7,
configs/charset/94_full.yaml:3.  Training-only keys are carried but never read by the engine.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Any, Dict, Tuple

CHARSET_94 = (
    "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    "!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"
)
CHARSET_36 = "0123456789abcdefghijklmnopqrstuvwxyz"


@dataclass
class ParseqConfig:
    # data
    charset_train: str = CHARSET_94
    charset_test: str = CHARSET_36
    max_label_length: int = 25
    img_size: Tuple[int, int] = (32, 128)      # (H, W)
    patch_size: Tuple[int, int] = (4, 8)       # (ph, pw)
    # architecture
    embed_dim: int = 384
    enc_num_heads: int = 6
    enc_mlp_ratio: int = 4
    enc_depth: int = 12
    dec_num_heads: int = 12
    dec_mlp_ratio: int = 4
    dec_depth: int = 1
    # decode mode
    decode_ar: bool = True
    refine_iters: int = 1
    # training-only (kept so that create_model(**yaml) round-trips; unused by the engine)
    batch_size: int = 384
    lr: float = 7e-4
    warmup_pct: float = 0.075
    weight_decay: float = 0.0
    perm_num: int = 6
    perm_forward: bool = True
    perm_mirrored: bool = True
    dropout: float = 0.1
    name: str = "parseq"
    # "parseq" | "vitstr" (strhub/models/vitstr: the same ViT with a class token + per-token head, no decoder;
    # embed_dim / enc_num_heads carry ViTSTR's embed_dim / num_heads, depth 12 and mlp_ratio 4 are fixed by
    # vitstr/system.py:50-59)
    arch: str = "parseq"
    extra: Dict[str, Any] = field(default_factory=dict)

    # ---- derived ----
    @property
    def num_tokens(self) -> int:          # EOS + charset + BOS + PAD
        return len(self.charset_train) + 3

    @property
    def num_classes(self) -> int:         # head never predicts BOS / PAD (model.py:62-63)
        return self.num_tokens - 2

    @property
    def grid(self) -> Tuple[int, int]:
        return (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])

    @property
    def num_patches(self) -> int:
        g = self.grid
        return g[0] * g[1]

    @property
    def enc_tokens(self) -> int:          # tokens per image inside the encoder (ViTSTR keeps timm's class token)
        return self.num_patches + (1 if self.arch == "vitstr" else 0)

    @property
    def patch_dim(self) -> int:
        return 3 * self.patch_size[0] * self.patch_size[1]

    @property
    def max_steps(self) -> int:           # +1 for EOS (model.py:110)
        return self.max_label_length + 1

    def to_kwargs(self) -> Dict[str, Any]:
        d = asdict(self)
        d.pop("extra")
        d["img_size"] = list(self.img_size)
        d["patch_size"] = list(self.patch_size)
        return d


PRESETS: Dict[str, Dict[str, Any]] = {
    # configs/model/parseq.yaml
    "parseq": dict(name="parseq"),
    # configs/experiment/parseq-tiny.yaml
    "parseq-tiny": dict(name="parseq-tiny", embed_dim=192, enc_num_heads=3, dec_num_heads=6),
    # configs/experiment/parseq-patch16-224.yaml
    "parseq-patch16-224": dict(name="parseq-patch16-224", img_size=(224, 224), patch_size=(16, 16)),
    # BASELINE.json configs[4]: ViT-B-width encoder stress config (not a reference experiment;
    # heads follow the D/64 (enc) and D/32 (dec) convention of the S / Ti configs)
    # configs/model/vitstr.yaml + configs/experiment/vitstr.yaml (32x128 crops, 4x8 patches, ViT-S width)
    "vitstr": dict(name="vitstr", arch="vitstr", lr=8.9e-4),
    "parseq-base-48x160": dict(name="parseq-base-48x160", embed_dim=768, enc_num_heads=12,
                               dec_num_heads=24, img_size=(48, 160)),
}


def make_config(experiment: str = "parseq", **overrides: Any) -> ParseqConfig:
    if experiment not in PRESETS:
        raise KeyError(experiment)
    kw = dict(PRESETS[experiment])
    known = set(ParseqConfig.__dataclass_fields__)
    extra = {}
    for k, v in overrides.items():
        if k in known:
            kw[k] = v
        else:
            extra[k] = v
    for k in ("img_size", "patch_size"):
        if k in kw:
            kw[k] = tuple(int(x) for x in kw[k])
    cfg = ParseqConfig(**kw)
    cfg.extra = extra
    return cfg
