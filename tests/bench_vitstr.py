"""ViTSTR-S timing (not a bench line): device time with CUDA events, bs=512 and bs=1, 32x128 / 4x8 (T = 128 + class token)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images

cfg = make_config("vitstr")
m = create_model("vitstr")
m.model.load_state_dict(init_state_dict(cfg, 0))
m = m.eval().to("cuda")
res = {}
for B, iters in ((512, 10), (1, 200)):
    x = synth_images(cfg, B, 3).cuda()
    with torch.inference_mode():
        for _ in range(3): out = m(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): out = m(x)
        b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    assert out.shape == (B, 26, 95) and torch.isfinite(out).all()
    res[f"ViTSTR-S bs={B}"] = {"ms": round(ms, 4), "images_per_s": round(B / ms * 1000, 1)}
    print(f"ViTSTR-S bs={B}: {ms:.3f} ms  {B / ms * 1000:.1f} img/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_vitstr.json", "w"), indent=1)
