"""`model.decode` / `model.head` / `model.text_embed` called on their own (".model must expose encode /
decode / head / text_embed / pos_queries"), against the CPU oracle's restatement of model.py:86-103 with explicit masks.
Sharp-attention weights: a wrong or ignored mask moves the logits far outside the tolerance."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_MAX, TOL_MEAN = 2.0e-2, 3.0e-3


def _setup(experiment="parseq", seed=12, sharp=4.0):
    from oracle.parseq_oracle import ParseqOracle
    from parseq_b200.config import make_config
    from parseq_b200.factory import create_model
    from parseq_b200.weights import init_state_dict
    cfg = make_config(experiment)
    sd = init_state_dict(cfg, seed, sharp=sharp)
    m = create_model(experiment)
    m.model.load_state_dict(sd)
    return cfg, sd, m.eval().to("cuda"), ParseqOracle(cfg, sd, "fp32")


def _inputs(cfg, oracle, B, J, seed):
    from parseq_b200.weights import synth_images
    g = torch.Generator().manual_seed(seed)
    x = synth_images(cfg, B, seed)
    memory = oracle.encode(x)                                   # fp32 [B, T, D]
    tgt = torch.randint(1, 95, (B, J), generator=g)
    tgt[:, 0] = cfg.num_tokens - 2                              # BOS
    return memory, tgt, g


def _check(model, oracle, memory, tgt, query, qmask, pmask, engine_kwargs):
    cfg = model.model.cfg
    B, J = tgt.shape
    q_or = query if query is not None else oracle.p["pos_queries"][:, :J].expand(B, -1, -1)
    ref = oracle._decode(tgt, memory, q_or, qmask, pmask)      # logits of head(decoder(...))
    with torch.inference_mode():
        out = model.model.decode(tgt.cuda(), memory.cuda(), **engine_kwargs)
        logits = model.model.head(out).cpu()
    assert out.shape == (B, q_or.shape[1], cfg.embed_dim) and out.dtype == torch.float32
    err = (logits - ref).abs()
    assert err.max().item() <= TOL_MAX and err.mean().item() <= TOL_MEAN, (err.max().item(), err.mean().item())
    return logits


@pytest.mark.parametrize("experiment", ["parseq", "parseq-tiny"])
def test_decode_causal_query_mask(experiment):
    cfg, sd, m, o = _setup(experiment)
    memory, tgt, _ = _inputs(cfg, o, 3, 10, 1)
    mask = torch.triu(torch.ones((10, 10), dtype=torch.bool), 1)               # model.py:117
    lm = _check(m, o, memory, tgt, None, mask, None, dict(tgt_query_mask=mask.cuda(), tgt_mask=mask.cuda()))
    # the mask matters: without it the result must differ by far more than the tolerance (sharp attention)
    with torch.inference_mode():
        lu = m.model.head(m.model.decode(tgt.cuda(), memory.cuda())).cpu()
    assert (lm - lu).abs().max().item() > 5 * TOL_MAX


def test_decode_cloze_and_padding_masks_additive_float():
    cfg, sd, m, o = _setup()
    L = 26
    memory, tgt, g = _inputs(cfg, o, 4, L, 2)
    qmask = torch.zeros((L, L), dtype=torch.bool)
    qmask[torch.arange(L - 1), torch.arange(1, L)] = True                       # model.py:157: only key i+1 hidden from query i
    pmask = torch.rand((4, L), generator=g) < 0.3
    pmask[:, 0] = False                                                         # BOS is never padding (model.py:163)
    addf = torch.zeros((L, L)).masked_fill(qmask, float("-inf"))               # additive float form of the same mask
    _check(m, o, memory, tgt, None, qmask, pmask, dict(tgt_query_mask=addf.cuda(), tgt_padding_mask=pmask.cuda()))


def test_decode_custom_queries_no_masks():
    cfg, sd, m, o = _setup()
    memory, tgt, g = _inputs(cfg, o, 2, 5, 3)
    query = torch.randn((2, 7, cfg.embed_dim), generator=g) * 0.05
    _check(m, o, memory, tgt, query, None, None, dict(tgt_query=query.cuda()))


def test_text_embed_and_head_modules():
    cfg, sd, m, o = _setup(sharp=0.0)
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(0, cfg.num_tokens, (5, 9), generator=g)
    with torch.inference_mode():
        e = m.model.text_embed(ids.cuda()).cpu()
    ref = math.sqrt(cfg.embed_dim) * sd["text_embed.embedding.weight"][ids]
    assert e.shape == (5, 9, cfg.embed_dim)
    assert (e - ref).abs().max().item() <= 1e-6
    x = torch.randn((3, 26, cfg.embed_dim), generator=g)
    with torch.inference_mode():
        lg = m.model.head(x.cuda()).cpu()
    ref = x.to(torch.bfloat16).float() @ sd["head.weight"].t() + sd["head.bias"]
    assert lg.shape == (3, 26, cfg.num_classes)
    assert (lg - ref).abs().max().item() <= 2e-3
    with pytest.raises(RuntimeError):
        m.model.head(x)                      # CPU tensor: no fallback
