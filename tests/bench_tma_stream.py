"""What can one CTA ingest through TMA?  Ring of 16 KB boxes, no compute, vs cluster size / ring depth / working set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
lib = load_library()
st = torch.cuda.current_stream().cuda_stream
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
def run(mb, cluster, ctas, nboxes, nslot, mode=0):
    buf = torch.empty(mb * 1024 * 1024 // 2, dtype=torch.bfloat16, device="cuda").normal_()
    def call():
        check(lib, lib.parseq_bench_tma_stream(buf.data_ptr(), buf.numel() * 2, cluster, ctas, nboxes, nslot, mode, sink.data_ptr(), st))
    call(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); call(); call(); call(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    per_sm = nboxes * 16384 / (ms * 1e-3) / 1e9
    return ms, per_sm, per_sm * ctas / 1e3
print("working set  cluster ctas slots mode            | ms/launch  GB/s per CTA  TB/s total   us per box")
names = {0: "cta-barrier, lane poll", 1: "cta-barrier, all poll", 2: "producer warp"}
for mb in (8, 1024):
    for cluster, ctas in ((1, 120), (8, 120), (1, 1)):
        for nslot in (3, 6, 12):
            for mode in (0, 2):
                ms, per, tot = run(mb, cluster, ctas, 2000, nslot, mode)
                print(f"{mb:6d} MB   {cluster:5d} {ctas:5d} {nslot:5d}  {names[mode]:22s} |  {ms:8.3f}   {per:9.1f}   {tot:8.2f}   {ms * 1000 / 2000:8.3f}")
