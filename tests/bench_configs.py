"""Timing of the other BASELINE.json configs (parity-test cases, not bench lines): device time, CUDA events."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images

def timeit(m, x, iters):
    with torch.inference_mode():
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): out = m(x)
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out

res = {}
for name, exp, B, ar, ri in [("C1 PARSeq-Ti bs=1 NAR refine0", "parseq-tiny", 1, False, 0),
                             ("PARSeq-Ti bs=512 AR+1", "parseq-tiny", 512, True, 1),
                             ("C2 PARSeq-S bs=512 AR+1", "parseq", 512, True, 1),
                             ("C4 PARSeq-S bs=1024 AR + 3 refine", "parseq", 1024, True, 3),
                             ("C5 ViT-B-width 48x160 (T=240, D=768) bs=256 AR+1", "parseq-base-48x160", 256, True, 1),
                             ("parseq-patch16-224 (T=196) bs=256 AR+1", "parseq-patch16-224", 256, True, 1),
                             ("PARSeq-S bs=1 AR+1", "parseq", 1, True, 1),
                             ("PARSeq-S bs=1 NAR + 3 refine (README.md:214-219 mode)", "parseq", 1, False, 3)]:
    cfg = make_config(exp)
    m = create_model(exp, decode_ar=ar, refine_iters=ri)
    m.model.load_state_dict(init_state_dict(cfg, 0))
    if B == 256:
        m.model.set_engine_option("max_batch", 256)
    m = m.eval().to("cuda")
    x = synth_images(cfg, B, 3).cuda()
    ms, out = timeit(m, x, 200 if B == 1 else 10)
    assert out.shape == (B, 26, 95) and torch.isfinite(out).all()
    if exp == "parseq-base-48x160":
        with torch.inference_mode():
            for _ in range(2): m.model.encode(x)
            torch.cuda.synchronize()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): m.model.encode(x)
            b2.record(); torch.cuda.synchronize()
        enc_ms = a.elapsed_time(b2) / 5
        res[name + " [encode only]"] = {"ms": round(enc_ms, 4), "images_per_s": round(B / enc_ms * 1000, 1)}
        print(f"{name + ' [encode only]':58s} {enc_ms:9.3f} ms  {B / enc_ms * 1000:10.1f} img/s", flush=True)
    res[name] = {"ms": round(ms, 4), "images_per_s": round(B / ms * 1000, 1)}
    print(f"{name:58s} {ms:9.3f} ms  {B / ms * 1000:10.1f} img/s", flush=True)
    del m
json.dump(res, open("gpurun_out/bench_configs.json", "w"), indent=1)
