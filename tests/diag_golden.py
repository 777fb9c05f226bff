"""Diagnostics: one golden case through the three AR-loop implementations (errors vs the fp32 reference)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images
name = sys.argv[1] if len(sys.argv) > 1 else "b48_sharp_ar1_b2"
blob = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".pt"), weights_only=False)
cfg = make_config(blob["experiment"])
sd = init_state_dict(cfg, blob["weight_seed"], sharp=blob.get("sharp", 0.0))
if blob.get("eos_bias"):
    sd["head.bias"] = sd["head.bias"].clone(); sd["head.bias"][0] += blob["eos_bias"]
m = create_model(blob["experiment"], decode_ar=blob["decode_ar"], refine_iters=blob["refine_iters"])
m.model.load_state_dict(sd)
m = m.eval().to("cuda")
x = synth_images(cfg, blob["batch"], blob["image_seed"])
L = 26
forced = torch.full((blob["batch"], L), 96, dtype=torch.int32); forced[:, : blob["ar_ids"].shape[1]] = blob["ar_ids"]
fr = torch.full((len(blob["refine_ctx"]), blob["batch"], L), 96, dtype=torch.int32)
for r, c in enumerate(blob["refine_ctx"]):
    fr[r, :, : c.shape[1]] = c; fr[r, :, c.shape[1]:] = 0
with torch.inference_mode():
    mem = m.model.encode(x.cuda()).cpu()[0]
    e = (mem - blob["memory0"]).abs()
    print(f"{name}: memory err max {e.max():.4f} mean {e.mean():.5f} (|mem| mean {blob['memory0'].abs().mean():.3f})")
    for impl in (2, 1, 0):
        m.model.set_engine_option("ar_kernel", impl)
        lg = m.model.forward(m.tokenizer, x.cuda(), blob["max_length"], forced_ids=forced, forced_refine=fr if len(blob["refine_ctx"]) else None).cpu()
        d = (lg - blob["logits"]).abs()
        print(f"  ar_kernel={impl}: logits err max {d.max():.4f} mean {d.mean():.5f} (sigma logits {blob['logits'].std():.3f})")
    m.model.decode_ar = True; m.model.refine_iters = 0
    outs = {}
    for impl in (2, 1, 0):
        m.model.set_engine_option("ar_kernel", impl)
        outs[impl] = m.model.forward(m.tokenizer, x.cuda(), 25, forced_ids=forced).cpu()
    from oracle.parseq_oracle import ParseqOracle
    o = ParseqOracle(cfg, sd, "fp32").forward(x, 25, True, 0, forced_ids=forced.long())
    for impl in (2, 1, 0):
        d = (outs[impl] - o.logits).abs()
        per_step = d.amax(dim=(0, 2))
        print(f"  AR-only ar_kernel={impl} vs fp32 oracle: max {d.max():.5f} mean {d.mean():.6f}; per-step max: "
              + " ".join(f"{v:.4f}" for v in per_step[:26:5].tolist()))
    dd = (outs[2] - outs[1]).abs().amax(dim=(0, 2))
    print("  |v2-v1| per-step max: " + " ".join(f"{v:.4f}" for v in dd.tolist()))
    print(f"  AR-only logits: |v2-v1| max {(outs[2]-outs[1]).abs().max():.5f}  |v2-eager| max {(outs[2]-outs[0]).abs().max():.5f}  |v1-eager| max {(outs[1]-outs[0]).abs().max():.5f}")
