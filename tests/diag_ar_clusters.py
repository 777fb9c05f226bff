"""How many 8-CTA clusters of the AR kernel are co-resident?  AR kernel time (engine timing category) vs batch and cluster count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images
cfg = make_config("parseq"); sd = init_state_dict(cfg, 0)
m = create_model("parseq", decode_ar=True, refine_iters=0); m.model.load_state_dict(sd)
m = m.eval().to("cuda")
eng = m.model.engine()
x = synth_images(cfg, 512, 1).cuda()
def ar_ms(B, clusters):
    eng.set_option("ar_clusters", clusters)
    with torch.inference_mode():
        m(x[:B], 25); m(x[:B], 25)
        torch.cuda.synchronize()
        eng.set_option("timing", 1)
        m(x[:B], 25)
        torch.cuda.synchronize()
        t = eng.get_timing()["dec_ar"]["ms"]
        eng.set_option("timing", 0)
    return t, eng.debug_int("ar_last_per"), eng.debug_int("ar_last_clusters")
print("sm_count", eng.debug_int("sm_count"))
for B, c in [(32, 1), (64, 2), (128, 4), (256, 8), (288, 9), (320, 10), (384, 12), (448, 14), (480, 15), (512, 16), (512, 0), (256, 16), (128, 16), (512, 32)]:
    t, per, ncl = ar_ms(B, c)
    print(f"B={B:4d} ar_clusters={c:2d} -> per={per:2d} clusters={ncl:2d}  AR kernel {t:7.3f} ms   (occupancy query mt1/mt2: {eng.debug_int('ar2_occupancy_mt1')}/{eng.debug_int('ar2_occupancy_mt2')})")
