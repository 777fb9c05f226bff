"""Seeded synthetic weights with the state_dict layout of the reference inner model.

Key names and shapes are those of `strhub.models.parseq.model.PARSeq.state_dict()`
(/root/reference/strhub/models/parseq/model.py:52-69, modules.py:31-43,145-161; timm ViT naming
for `encoder.*`), so a released `parseq-*.pt` file and a tensor dict produced here are
interchangeable inputs of `PARSeq.model.load_state_dict`.

Distributions follow the reference initialisers (strhub/models/utils.py:107-125, model.py:70-71,
timm ViT: trunc-normal(0.02) linears / pos_embed) with two deliberate differences, both so that
parity tests have teeth:
  * `perturb=True` gives biases and LayerNorm affine parameters non-trivial values (the reference
    initialises them to 0 / 1, which would let a dropped bias or gamma go unnoticed);
  * `bf16_exact=True` rounds the GEMM weight matrices (and nothing else) to bf16-representable
    values, so the engine's one-time bf16 weight packing is lossless and the fp32 reference, the
    oracle and the engine all see identical weights.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Dict

import torch

from .config import ParseqConfig


def _tn(gen: torch.Generator, shape, std: float = 0.02) -> torch.Tensor:
    t = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0, generator=gen)
    return t


def gemm_weight_keys(cfg: ParseqConfig):
    """state_dict keys the engine stores in bf16 (tensor-core operands)."""
    if getattr(cfg, "arch", "parseq") == "vitstr":
        keys = ["patch_embed.proj.weight"]
        for i in range(cfg.enc_depth):
            p = f"blocks.{i}."
            keys += [p + "attn.qkv.weight", p + "attn.proj.weight", p + "mlp.fc1.weight", p + "mlp.fc2.weight"]
        return keys + ["head.weight"]
    keys = ["encoder.patch_embed.proj.weight"]
    for i in range(cfg.enc_depth):
        p = f"encoder.blocks.{i}."
        keys += [p + "attn.qkv.weight", p + "attn.proj.weight", p + "mlp.fc1.weight", p + "mlp.fc2.weight"]
    for i in range(cfg.dec_depth):
        p = f"decoder.layers.{i}."
        keys += [p + "self_attn.in_proj_weight", p + "self_attn.out_proj.weight",
                 p + "cross_attn.in_proj_weight", p + "cross_attn.out_proj.weight",
                 p + "linear1.weight", p + "linear2.weight"]
    keys += ["head.weight"]
    return keys


def init_state_dict(cfg: ParseqConfig, seed: int = 0, perturb: bool = True,
                    bf16_exact: bool = True, sharp: float = 0.0) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    D = cfg.embed_dim
    T = cfg.num_patches
    ph, pw = cfg.patch_size
    Me = D * cfg.enc_mlp_ratio
    Md = D * cfg.dec_mlp_ratio
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def bias(n):
        return _tn(g, (n,), 0.02) if perturb else torch.zeros(n)

    def ln(prefix):
        if perturb:
            sd[prefix + ".weight"] = 1.0 + _tn(g, (D,), 0.1)
            sd[prefix + ".bias"] = _tn(g, (D,), 0.05)
        else:
            sd[prefix + ".weight"] = torch.ones(D)
            sd[prefix + ".bias"] = torch.zeros(D)

    if getattr(cfg, "arch", "parseq") == "vitstr":
        return _init_vitstr(cfg, g, sd, bias, ln, perturb, bf16_exact)
    # ---- encoder (timm ViT names) ----
    sd["encoder.pos_embed"] = _tn(g, (1, T, D))
    fan_in = 3 * ph * pw
    bound = (1.0 / fan_in) ** 0.5          # nn.Conv2d default init scale (kaiming-uniform a=sqrt(5))
    sd["encoder.patch_embed.proj.weight"] = (torch.rand((D, 3, ph, pw), generator=g) * 2 - 1) * bound
    sd["encoder.patch_embed.proj.bias"] = (torch.rand((D,), generator=g) * 2 - 1) * bound
    for i in range(cfg.enc_depth):
        p = f"encoder.blocks.{i}."
        ln(p + "norm1")
        sd[p + "attn.qkv.weight"] = _tn(g, (3 * D, D))
        sd[p + "attn.qkv.bias"] = bias(3 * D)
        sd[p + "attn.proj.weight"] = _tn(g, (D, D))
        sd[p + "attn.proj.bias"] = bias(D)
        ln(p + "norm2")
        sd[p + "mlp.fc1.weight"] = _tn(g, (Me, D))
        sd[p + "mlp.fc1.bias"] = bias(Me)
        sd[p + "mlp.fc2.weight"] = _tn(g, (D, Me))
        sd[p + "mlp.fc2.bias"] = bias(D)
    ln("encoder.norm")
    # ---- decoder ----
    for i in range(cfg.dec_depth):
        p = f"decoder.layers.{i}."
        for att in ("self_attn", "cross_attn"):
            sd[p + att + ".in_proj_weight"] = _tn(g, (3 * D, D))
            sd[p + att + ".in_proj_bias"] = bias(3 * D)
            sd[p + att + ".out_proj.weight"] = _tn(g, (D, D))
            sd[p + att + ".out_proj.bias"] = bias(D)
        sd[p + "linear1.weight"] = _tn(g, (Md, D))
        sd[p + "linear1.bias"] = bias(Md)
        sd[p + "linear2.weight"] = _tn(g, (D, Md))
        sd[p + "linear2.bias"] = bias(D)
        for n in ("norm1", "norm2", "norm_q", "norm_c"):
            ln(p + n)
    ln("decoder.norm")
    sd["head.weight"] = _tn(g, (cfg.num_classes, D))
    sd["head.bias"] = bias(cfg.num_classes)
    sd["text_embed.embedding.weight"] = _tn(g, (cfg.num_tokens, D))
    sd["pos_queries"] = _tn(g, (1, cfg.max_steps, D))

    if sharp:
        _sharpen(cfg, sd, float(sharp))
    if bf16_exact:
        for k in gemm_weight_keys(cfg):
            sd[k] = sd[k].to(torch.bfloat16).to(torch.float32)
    return sd


def _sharpen(cfg: ParseqConfig, sd, s: float):
    """"Sharp" synthetic weights: the query and key rows of every attention in-projection (and their biases) are
    multiplied by `s`, i.e. every pre-softmax score by s^2.  With the reference's trunc-normal(0.02) init all scores
    have sigma ~ 0.15 and every softmax is nearly uniform, which hides whole classes of attention bugs (a wrong or
    missing q, mask or scale moves the logits by less than the bf16 tolerance); s = 4 gives sigma ~ 2.4 (peaked
    rows).  A power of two keeps bf16-exact weights bf16-exact."""
    D = cfg.embed_dim
    keys = [f"encoder.blocks.{i}.attn.qkv" for i in range(cfg.enc_depth)]
    for i in range(cfg.dec_depth):
        keys += [f"decoder.layers.{i}.self_attn.in_proj", f"decoder.layers.{i}.cross_attn.in_proj"]
    for k in keys:
        w = k + (".weight" if k.endswith("qkv") else "_weight")
        b = k + (".bias" if k.endswith("qkv") else "_bias")
        sd[w] = sd[w].clone()
        sd[b] = sd[b].clone()
        sd[w][: 2 * D] *= s
        sd[b][: 2 * D] *= s


def _init_vitstr(cfg, g, sd, bias, ln, perturb, bf16_exact):
    """state_dict of strhub.models.vitstr.model.ViTSTR == timm VisionTransformer(class_token=True, num_classes=C)
    (vitstr/system.py:50-61): no "encoder." prefix, `cls_token`, `pos_embed` over T+1 tokens, `head`."""
    D, T = cfg.embed_dim, cfg.num_patches
    ph, pw = cfg.patch_size
    Me = D * cfg.enc_mlp_ratio
    sd["cls_token"] = _tn(g, (1, 1, D)) if perturb else torch.zeros(1, 1, D)   # timm: normal(std=1e-6)
    sd["pos_embed"] = _tn(g, (1, T + 1, D))
    bound = (1.0 / (3 * ph * pw)) ** 0.5
    sd["patch_embed.proj.weight"] = (torch.rand((D, 3, ph, pw), generator=g) * 2 - 1) * bound
    sd["patch_embed.proj.bias"] = (torch.rand((D,), generator=g) * 2 - 1) * bound
    for i in range(cfg.enc_depth):
        p = f"blocks.{i}."
        ln(p + "norm1")
        sd[p + "attn.qkv.weight"] = _tn(g, (3 * D, D))
        sd[p + "attn.qkv.bias"] = bias(3 * D)
        sd[p + "attn.proj.weight"] = _tn(g, (D, D))
        sd[p + "attn.proj.bias"] = bias(D)
        ln(p + "norm2")
        sd[p + "mlp.fc1.weight"] = _tn(g, (Me, D))
        sd[p + "mlp.fc1.bias"] = bias(Me)
        sd[p + "mlp.fc2.weight"] = _tn(g, (D, Me))
        sd[p + "mlp.fc2.bias"] = bias(D)
    ln("norm")
    sd["head.weight"] = _tn(g, (cfg.num_classes, D))      # vitstr/system.py:61 -> init_weights: trunc-normal(0.02)
    sd["head.bias"] = bias(cfg.num_classes)
    if bf16_exact:
        for k in gemm_weight_keys(cfg):
            sd[k] = sd[k].to(torch.bfloat16).to(torch.float32)
    return sd


def synth_images(cfg: ParseqConfig, batch: int, seed: int = 0, bf16_exact: bool = True) -> torch.Tensor:
    """Seeded crops in [-1, 1) (the range `T.Normalize(0.5, 0.5)` produces, data/module.py:77-81)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1_000_003 + seed)
    x = torch.rand((batch, 3, cfg.img_size[0], cfg.img_size[1]), generator=g) * 2 - 1
    if bf16_exact:
        x = x.to(torch.bfloat16).to(torch.float32)
    return x


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def count_params(sd: Dict[str, torch.Tensor]) -> int:
    return sum(int(v.numel()) for v in sd.values())
