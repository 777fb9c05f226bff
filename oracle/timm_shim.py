"""Minimal stand-in for the three timm names the reference imports, so that
/root/reference/strhub/models/parseq/{model,modules}.py can be imported UNMODIFIED in this
container (timm==0.9.16 is pinned by requirements/core.txt:32 but not installed, no network).

TEST INFRASTRUCTURE ONLY (used by oracle/reference_loader.py to generate golden vectors).

The reference touches: `timm.models.vision_transformer.{VisionTransformer, PatchEmbed}`
(modules.py:24) and `timm.models.helpers.named_apply` (model.py:23).  This file restates the
timm-0.9.16 ViT forward for exactly the constructor arguments the reference passes
(modules.py:145-161: num_classes=0, global_pool='', class_token=False, qkv_bias=True, all drop
rates 0) with timm's parameter names, so `state_dict()` keys match released PARSeq weights, and
for the ctor of strhub/models/vitstr/system.py:50-59 (timm defaults class_token=True,
global_pool='token', num_classes=len(tokenizer)-2: `cls_token`, `pos_embed` [1, T+1, D], `head`).
It is my own restatement of published semantics, not timm code => encoder "parity unpinned".
"""
from __future__ import annotations

import sys
import types

import torch
from torch import nn
import torch.nn.functional as F


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, **_):
        super().__init__()
        self.img_size = _pair(img_size)
        self.patch_size = _pair(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)

    def forward(self, x):
        assert tuple(x.shape[-2:]) == self.img_size, "input size does not match the model"
        return self.proj(x).flatten(2).transpose(1, 2)


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        return self.proj(o.transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, global_pool='token',
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True, class_token=True,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, **_):
        super().__init__()
        assert (num_classes == 0 and global_pool == '' and not class_token) or \
               (num_classes > 0 and global_pool == 'token' and class_token), "shim covers the PARSeq and ViTSTR ctors only"
        assert drop_rate == attn_drop_rate == drop_path_rate == 0.0
        self.embed_dim = embed_dim
        self.num_classes = num_classes
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                       embed_dim=embed_dim)
        # timm 0.9.16 VisionTransformer.__init__: cls_token [1,1,D]; pos_embed covers the prefix token too
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if class_token else None
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + (1 if class_token else 0), embed_dim))
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        if self.cls_token is not None:
            nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def forward_features(self, x):
        x = self.patch_embed(x)
        if self.cls_token is not None:       # timm _pos_embed (no_embed_class=False): concat, then add
            x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1)
        x = x + self.pos_embed
        return self.norm(self.blocks(x))

    def forward(self, x):
        x = self.forward_features(x)
        if self.num_classes > 0:             # global_pool='token'
            return self.head(x[:, 0])
        return x


def named_apply(fn, module, name='', depth_first=True, include_root=False):
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        full = '.'.join((name, child_name)) if name else child_name
        named_apply(fn=fn, module=child, name=full, depth_first=depth_first, include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module


def install():
    if 'timm' in sys.modules and not getattr(sys.modules['timm'], '_parseq_b200_shim', False):
        return  # a real timm is importable: use it
    timm = types.ModuleType('timm'); timm._parseq_b200_shim = True
    models = types.ModuleType('timm.models')
    vt = types.ModuleType('timm.models.vision_transformer')
    helpers = types.ModuleType('timm.models.helpers')
    vt.VisionTransformer = VisionTransformer
    vt.PatchEmbed = PatchEmbed
    helpers.named_apply = named_apply
    timm.models = models
    models.vision_transformer = vt
    models.helpers = helpers
    sys.modules.update({'timm': timm, 'timm.models': models,
                        'timm.models.vision_transformer': vt, 'timm.models.helpers': helpers})
