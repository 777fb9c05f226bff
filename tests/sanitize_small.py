"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck) on the final build: PARSeq-Ti and -S, all
decode modes; the cluster AR kernel in both cluster sizes (8, and 6 forced) and the grid-barrier AR kernel; the fused
residual-GEMM + LayerNorm kernels forced; the tcgen05 attention; PARSeq.decode with masks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images

for exp in ("parseq-tiny", "parseq"):
    cfg = make_config(exp)
    sd = init_state_dict(cfg, 0)
    for ar, ri, ml, impl, cs, B in [(True, 1, None, 2, 0, 3), (True, 0, 4, 2, 6, 19), (False, 2, None, 2, 0, 3), (True, 0, 4, 1, 0, 3)]:
        m = create_model(exp, decode_ar=ar, refine_iters=ri)
        m.model.load_state_dict(sd)
        m.model.set_engine_option("use_graph", 0)
        m.model.set_engine_option("ar_kernel", impl)
        m.model.set_engine_option("ar_cluster_size", cs)
        m.model.set_engine_option("fuse_ln", 7 if ar else 0)     # AR modes: fused residual-GEMM + LayerNorm kernels forced
        m = m.eval().to("cuda")
        x = synth_images(cfg, B, 1).cuda()
        with torch.inference_mode():
            out = m(x, ml)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        print("ok:", exp, "ar", ar, "refine", ri, "max_length", ml, "ar_kernel", impl, "cluster", cs, tuple(out.shape), flush=True)
    # decode API with masks
    m = create_model(exp); m.model.load_state_dict(sd); m = m.eval().to("cuda")
    x = synth_images(cfg, 2, 3).cuda()
    with torch.inference_mode():
        mem = m.model.encode(x)
        tgt = torch.randint(1, 90, (2, 9), device="cuda"); tgt[:, 0] = 95
        qm = torch.triu(torch.ones((9, 9), dtype=torch.bool, device="cuda"), 1)
        out = m.model.head(m.model.decode(tgt, mem, tgt_query_mask=qm))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    print("ok:", exp, "decode + head", tuple(out.shape), flush=True)
print("sanitize_small done")
