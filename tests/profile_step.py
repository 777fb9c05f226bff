"""Profiling driver (not a test): one warm forward, then one profiled forward between cudaProfilerStart/Stop.
    ncu --profile-from-start off ... python tests/profile_step.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = make_config("parseq")
m = create_model("parseq", decode_ar=True, refine_iters=1)
m.model.load_state_dict(init_state_dict(cfg, 0))
if os.environ.get("PQ_FUSE_MLP"):        # opt-in one-kernel MLP + LayerNorm (mlp_ln.cuh), to capture its traffic inside a step
    m.model.set_engine_option("fuse_mlp", 1)
    if os.environ.get("PQ_MLP_CTA_GROUP"):
        m.model.set_engine_option("mlp_cta_group", int(os.environ["PQ_MLP_CTA_GROUP"]))
m = m.eval().to("cuda")
x = synth_images(cfg, B, 1).cuda()
with torch.inference_mode():
    m(x); m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one forward of", B, "images; launches/forward:", m.model.engine().launches // 3)
