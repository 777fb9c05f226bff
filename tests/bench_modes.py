"""Timing decomposition (not a test): graph-mode device time of forward() under different decode modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images

cfg = make_config("parseq")
sd = init_state_dict(cfg, 0)
B = 512
x = synth_images(cfg, B, 1).cuda()

def timeit(m, ml, iters=10):
    with torch.inference_mode():
        for _ in range(3): m(x, ml)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): m(x, ml)
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

for dec_chunk in (128,):
    for pdl in (1,):
        res = {}
        for name, ar, ri, ml in [("nar0", False, 0, None), ("ar0_len0(1 step)", True, 0, 0), ("ar0", True, 0, 25), ("ar1", True, 1, None), ("nar1", False, 1, None)]:
            m = create_model("parseq", decode_ar=ar, refine_iters=ri)
            m.model.load_state_dict(sd)
            m.model.set_engine_option("dec_chunk", dec_chunk)
            m.model.set_engine_option("pdl", pdl)
            m = m.eval().to("cuda")
            res[name] = timeit(m, ml)
            del m
        print(f"dec_chunk={dec_chunk} pdl={pdl}: " + "  ".join(f"{k}={v:.3f}ms" for k, v in res.items()),
              f" | per AR step ~ {(res['ar0'] - res['ar0_len0(1 step)']) / 25 * 1000:.1f} us; refine pass ~ {(res['ar1'] - res['ar0']) * 1000:.0f} us")
