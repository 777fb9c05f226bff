"""CPU: independent third-party pin of the ViT arithmetic behind every golden.

The reference's encoder is `timm==0.9.16` `VisionTransformer.forward_features` (strhub/models/parseq/modules.py:128-165),
which is neither vendored in /root/reference nor installable here; `oracle/timm_shim.py` and `ParseqOracle.encode` are
restatements.  torchvision 0.26 (installed in this image) ships an independently written pre-LN ViT encoder
(`torchvision.models.vision_transformer.{EncoderBlock, Encoder, VisionTransformer}`: LayerNorm eps 1e-6,
`nn.MultiheadAttention` with packed in-projection, erf-GELU MLP, class token + learned position embedding).  Mapping
  blocks.i.norm1 -> ln_1,  attn.qkv -> self_attention.in_proj_{weight,bias},  attn.proj -> self_attention.out_proj,
  blocks.i.norm2 -> ln_2,  mlp.fc1 / mlp.fc2 -> mlp.0 / mlp.3,  pos_embed -> pos_embedding,  norm -> ln
the shim, the oracle and torchvision must agree to fp32 round-off over all 12 blocks.  What stays pinned only by the
reference's constructor arguments (modules.py:145-161) is WHICH ViT is built (no class token, no pooling, eps, widths).
"""
import pytest
import torch

tv = pytest.importorskip("torchvision.models.vision_transformer")

TOL = 1e-5


def _tv_encoder(cfg, sd, prefix, seq_len):
    D = cfg.embed_dim
    enc = tv.Encoder(seq_length=seq_len, num_layers=cfg.enc_depth, num_heads=cfg.enc_num_heads, hidden_dim=D,
                     mlp_dim=D * cfg.enc_mlp_ratio, dropout=0.0, attention_dropout=0.0).eval()
    m = {"pos_embedding": sd[prefix + "pos_embed"], "ln.weight": sd[prefix + "norm.weight"], "ln.bias": sd[prefix + "norm.bias"]}
    for i in range(cfg.enc_depth):
        s, d = f"{prefix}blocks.{i}.", f"layers.encoder_layer_{i}."
        m[d + "ln_1.weight"], m[d + "ln_1.bias"] = sd[s + "norm1.weight"], sd[s + "norm1.bias"]
        m[d + "self_attention.in_proj_weight"] = sd[s + "attn.qkv.weight"]
        m[d + "self_attention.in_proj_bias"] = sd[s + "attn.qkv.bias"]
        m[d + "self_attention.out_proj.weight"] = sd[s + "attn.proj.weight"]
        m[d + "self_attention.out_proj.bias"] = sd[s + "attn.proj.bias"]
        m[d + "ln_2.weight"], m[d + "ln_2.bias"] = sd[s + "norm2.weight"], sd[s + "norm2.bias"]
        m[d + "mlp.0.weight"], m[d + "mlp.0.bias"] = sd[s + "mlp.fc1.weight"], sd[s + "mlp.fc1.bias"]
        m[d + "mlp.3.weight"], m[d + "mlp.3.bias"] = sd[s + "mlp.fc2.weight"], sd[s + "mlp.fc2.bias"]
    missing, unexpected = enc.load_state_dict(m, strict=True)
    return enc


def _patch_tokens(cfg, sd, prefix, x):
    """Conv2d(k = s = patch) + flatten(2).transpose(1, 2) with torch's own conv (timm PatchEmbed, torchvision
    `_process_input`)."""
    w, b = sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"]
    return torch.nn.functional.conv2d(x, w, b, stride=cfg.patch_size).flatten(2).transpose(1, 2)


@pytest.mark.parametrize("experiment,seed", [("parseq", 0), ("parseq-tiny", 2)])
def test_oracle_and_shim_encoder_match_torchvision(experiment, seed):
    from oracle import timm_shim
    from oracle.parseq_oracle import ParseqOracle
    from parseq_b200.config import make_config
    from parseq_b200.weights import init_state_dict, synth_images
    cfg = make_config(experiment)
    sd = init_state_dict(cfg, seed)
    x = synth_images(cfg, 3, 21)
    with torch.no_grad():
        ref = _tv_encoder(cfg, sd, "encoder.", cfg.num_patches)(_patch_tokens(cfg, sd, "encoder.", x))
    # (1) the oracle's restatement
    mem = ParseqOracle(cfg, sd, "fp32").encode(x)
    assert mem.shape == ref.shape == (3, cfg.num_patches, cfg.embed_dim)
    assert (mem - ref).abs().max().item() <= TOL
    # (2) the timm stand-in every golden was generated with (ctor args of modules.py:145-161)
    vit = timm_shim.VisionTransformer(img_size=list(cfg.img_size), patch_size=list(cfg.patch_size), embed_dim=cfg.embed_dim,
                                      depth=cfg.enc_depth, num_heads=cfg.enc_num_heads, mlp_ratio=cfg.enc_mlp_ratio,
                                      qkv_bias=True, num_classes=0, global_pool='', class_token=False).eval()
    vit.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    with torch.no_grad():
        shim = vit.forward_features(x)
    assert (shim - ref).abs().max().item() <= TOL


def test_single_block_matches_torchvision_encoder_block():
    """Block-level pin (residual placement, pre-LN order, head split of the packed qkv projection)."""
    from oracle import timm_shim
    torch.manual_seed(5)
    D, H = 384, 6
    blk = timm_shim._Block(D, H, 4, True).eval()
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.05)
    ref = tv.EncoderBlock(H, D, 4 * D, 0.0, 0.0).eval()
    ref.load_state_dict({
        "ln_1.weight": blk.norm1.weight, "ln_1.bias": blk.norm1.bias,
        "self_attention.in_proj_weight": blk.attn.qkv.weight, "self_attention.in_proj_bias": blk.attn.qkv.bias,
        "self_attention.out_proj.weight": blk.attn.proj.weight, "self_attention.out_proj.bias": blk.attn.proj.bias,
        "ln_2.weight": blk.norm2.weight, "ln_2.bias": blk.norm2.bias,
        "mlp.0.weight": blk.mlp.fc1.weight, "mlp.0.bias": blk.mlp.fc1.bias,
        "mlp.3.weight": blk.mlp.fc2.weight, "mlp.3.bias": blk.mlp.fc2.bias}, strict=True)
    x = torch.randn(2, 128, D)
    with torch.no_grad():
        assert (blk(x) - ref(x)).abs().max().item() <= TOL


def test_vitstr_class_token_path_matches_torchvision():
    """ViTSTR keeps timm's class token (vitstr/model.py:14-28): concat the token, add pos_embed over T + 1 tokens, blocks,
    norm.  torchvision's full VisionTransformer does the same for square geometries (configs/model/vitstr.yaml: 224 / 16)."""
    from oracle.vitstr_oracle import VitstrOracle
    from parseq_b200.config import make_config
    from parseq_b200.weights import init_state_dict, synth_images
    cfg = make_config("vitstr", img_size=(224, 224), patch_size=(16, 16))
    sd = init_state_dict(cfg, 21)
    x = synth_images(cfg, 1, 33)
    enc = _tv_encoder(cfg, sd, "", cfg.num_patches + 1)
    with torch.no_grad():
        tok = _patch_tokens(cfg, sd, "", x)
        ref = enc(torch.cat([sd["cls_token"].expand(1, -1, -1), tok], dim=1))
    feats = VitstrOracle(cfg, sd, "fp32").features(x)
    assert (feats - ref).abs().max().item() <= TOL
