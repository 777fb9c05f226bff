"""Phase time stamps of the persistent AR kernels (engine option "ar_prof"): python tests/prof_ar.py [batch] [ar_kernel]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = make_config("parseq"); sd = init_state_dict(cfg, 0)
m = create_model("parseq", decode_ar=True, refine_iters=0); m.model.load_state_dict(sd)
m.model.set_engine_option("ar_kernel", impl)
m.model.set_engine_option("ar_prof", 1)
m = m.eval().to("cuda")
x = synth_images(cfg, B, 1).cuda()
with torch.inference_mode():
    for _ in range(3): m(x, 25)
torch.cuda.synchronize()
prof = m.model.engine().get_ar_profile()
if impl == 1:
    names = ["P1 self", "bar", "P2 oproj", "bar", "P3 ln+q", "bar", "P4 cross", "bar", "P5 oproj", "bar", "P6 ln+l1", "bar", "P7 l2", "bar", "P8 head"]
else:
    names = ["P1 self", "sync", "P2 oproj+st", "sync+ln1", "sync+P3 q", "sync", "P4 cross", "sync", "P5 oproj+st", "sync+ln2+sync", "P6 l1", "P7 l2", "sync+red+st", "sync+ln3+sync", "P8 head"]
print(f"batch {B} ar_kernel {impl}")
for step in (0, 1, 12, 13, 25):
    t = prof[step]
    print(f"step {step}: " + "  ".join(f"{names[k]}={(t[k+1]-t[k])/1000:.1f}" for k in range(15)), f" | total {(t[15]-t[0])/1000:.1f} us")
print(f"whole loop: {(prof[25][15]-prof[0][0])/1000:.1f} us")
if impl == 2:
    t = prof[26]
    if t[0]:
        print("P4 detail (step 1, cluster 0 rank 0), us from P4 start; per owned image: K loop done, softmax done, V loop + broadcast done:")
        print("  " + "  ".join(f"{(t[k]-t[0])/1000:.1f}" for k in range(1, 13) if t[k]))
    print("cluster size", m.model.engine().debug_int("ar_last_cluster_size"), "rows/cluster", m.model.engine().debug_int("ar_last_per"),
          "clusters", m.model.engine().debug_int("ar_last_clusters"))
