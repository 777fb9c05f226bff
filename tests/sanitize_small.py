"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): PARSeq-Ti and -S, B=3, all decode modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images

for exp in ("parseq-tiny", "parseq"):
    cfg = make_config(exp)
    sd = init_state_dict(cfg, 0)
    for ar, ri, ml in [(True, 1, None), (False, 2, None), (True, 0, 4)]:
        m = create_model(exp, decode_ar=ar, refine_iters=ri)
        m.model.load_state_dict(sd)
        m.model.set_engine_option("use_graph", 0)
        m.model.set_engine_option("fuse_ln", 7 if ar else 0)     # AR modes: fused residual-GEMM + LayerNorm kernels forced
        m = m.eval().to("cuda")
        x = synth_images(cfg, 3, 1).cuda()
        with torch.inference_mode():
            out = m(x, ml)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        print(exp, ar, ri, ml, tuple(out.shape), flush=True)
print("sanitize_small done")
