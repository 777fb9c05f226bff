"""CPU: ViTSTR — the oracle against golden outputs of the reference's own
strhub.models.vitstr.model.ViTSTR (tests/golden/vitstr_*.pt, oracle/make_golden.py vitstr), and the host-side mirror
of strhub.models.vitstr.system.ViTSTR / hubconf.vitstr."""
import glob
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "vitstr_*.pt")))


def _cfg_sd(blob):
    from parseq_b200.config import make_config
    from parseq_b200.weights import init_state_dict
    cfg = make_config("vitstr", **blob["overrides"])
    return cfg, init_state_dict(cfg, blob["weight_seed"])


def test_golden_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[:-3])
def test_oracle_fp32_matches_reference_golden(path):
    from oracle.vitstr_oracle import VitstrOracle
    from parseq_b200.weights import synth_images, state_dict_digest
    blob = torch.load(path, weights_only=False)
    cfg, sd = _cfg_sd(blob)
    assert state_dict_digest(sd) == blob["sd_digest"], "synthetic weight generator changed: regenerate tests/golden"
    x = synth_images(cfg, blob["batch"], blob["image_seed"])
    o = VitstrOracle(cfg, sd, "fp32")
    logits = o.system_forward(x, blob["max_length"])
    assert logits.shape == blob["logits"].shape              # [B, min(max_length, 25) + 1, 95] (vitstr/system.py:66-70)
    assert (logits - blob["logits"]).abs().max().item() < 2e-5
    assert (o.features(x[:1])[0] - blob["features0"]).abs().max().item() < 2e-5
    assert torch.equal(logits.argmax(-1), blob["logits"].argmax(-1))


def test_model_forward_keeps_the_class_token_row_and_system_drops_it():
    """vitstr/model.py:21 keeps tokens [0, seqlen); vitstr/system.py:70 drops token 0."""
    from oracle.vitstr_oracle import VitstrOracle
    from parseq_b200.config import make_config
    from parseq_b200.weights import init_state_dict, synth_images
    cfg = make_config("vitstr")
    o = VitstrOracle(cfg, init_state_dict(cfg, 3), "fp32")
    x = synth_images(cfg, 2, 1)
    full = o.model_forward(x, 9)
    assert full.shape == (2, 9, 95)
    assert torch.equal(o.system_forward(x, 7), full[:, 1:])
    assert o.system_forward(x, 99).shape == (2, 26, 95)      # clamp to max_label_length


def test_bf16_mode_is_a_small_perturbation_of_fp32():
    from oracle.vitstr_oracle import VitstrOracle
    from parseq_b200.config import make_config
    from parseq_b200.weights import init_state_dict, synth_images
    cfg = make_config("vitstr")
    sd = init_state_dict(cfg, 4)
    x = synth_images(cfg, 2, 2)
    a = VitstrOracle(cfg, sd, "fp32").system_forward(x)
    b = VitstrOracle(cfg, sd, "bf16").system_forward(x)
    err = (a - b).abs()
    assert 1e-5 < err.max().item() < 3e-2 and err.mean().item() < 4e-3


def test_state_dict_layout_is_the_timm_vit_with_class_token():
    from parseq_b200.config import make_config
    from parseq_b200.weights import init_state_dict, count_params
    cfg = make_config("vitstr")
    sd = init_state_dict(cfg, 0)
    assert cfg.enc_tokens == 129 and cfg.num_patches == 128
    assert tuple(sd["cls_token"].shape) == (1, 1, 384) and tuple(sd["pos_embed"].shape) == (1, 129, 384)
    assert tuple(sd["head.weight"].shape) == (95, 384)       # len(tokenizer) - 2 (vitstr/system.py:58)
    assert not any(k.startswith(("encoder.", "decoder.")) for k in sd)
    # PARSeq-S encoder (21,380,736: README.md:220-226) + cls_token 384 + one more pos_embed row 384 + head 36,575
    assert count_params(sd) == 21_380_736 + 384 + 384 + 36_575


def test_host_mirror_of_the_reference_surface():
    import hubconf
    from strhub.models.utils import create_model, InvalidModelError
    from strhub.models.vitstr.system import ViTSTR
    from strhub.models.vitstr.model import ViTSTR as Inner
    m = hubconf.vitstr()
    assert isinstance(m, ViTSTR) and isinstance(m.model, Inner)
    assert m.hparams.img_size == [32, 128] and m.hparams.patch_size == [4, 8]      # configs/experiment/vitstr.yaml
    assert m.hparams.embed_dim == 384 and m.hparams.num_heads == 6 and m.hparams.lr == 8.9e-4
    assert m.max_label_length == 25 and len(m.tokenizer) == 97
    big = create_model("vitstr", img_size=[224, 224], patch_size=[16, 16])        # configs/model/vitstr.yaml geometry
    assert tuple(big.model.pos_embed.shape) == (1, 197, 384)
    with pytest.raises(InvalidModelError):
        create_model("crnn")
    # released ViTSTR weights are saved from the system ('model.' prefix, strhub/models/utils.py:80-82)
    from parseq_b200.weights import init_state_dict
    sd = init_state_dict(m.model.cfg, 1)
    m.load_state_dict({"model." + k: v for k, v in sd.items()})
    assert torch.equal(m.model.head.weight, sd["head.weight"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 128))
