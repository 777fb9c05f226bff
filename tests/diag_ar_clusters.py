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
def ar_ms(B, clusters, cs=0):
    eng.set_option("ar_clusters", clusters)
    eng.set_option("ar_cluster_size", cs)
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
for B, c, cs in [(480, 0, 8), (512, 0, 8), (512, 0, 6), (512, 0, 0), (1, 0, 0), (16, 0, 0), (64, 0, 0), (240, 0, 0), (256, 0, 6), (736, 0, 6)]:
    if B > 512:
        continue
    t, per, ncl = ar_ms(B, c, cs)
    print(f"B={B:4d} forced cluster size {cs} -> cluster size {eng.debug_int('ar_last_cluster_size')} per={per:2d} clusters={ncl:2d}  AR kernel {t:7.3f} ms")
print("occupancy query (clusters): cs8 mt1/mt2", eng.debug_int("ar2_occupancy_mt1_cs8"), eng.debug_int("ar2_occupancy_mt2_cs8"),
      " cs6 mt1/mt2", eng.debug_int("ar2_occupancy_mt1_cs6"), eng.debug_int("ar2_occupancy_mt2_cs6"))
