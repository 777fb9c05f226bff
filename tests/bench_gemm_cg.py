"""cta_group 1 vs 2 over the GEMM shapes of the supported widths (M = rows of a 256- / 512-image batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
lib = load_library()
st = torch.cuda.current_stream().cuda_stream
def run(M, N, K, mode, inplace, cg, bn=0, iters=20):
    check(lib, lib.parseq_set_option(None, b"cta_group", cg))
    check(lib, lib.parseq_set_option(None, b"block_n", bn))
    A = torch.randn((M, K), device="cuda").bfloat16()
    W = (torch.randn((N, K), device="cuda") * 0.02).bfloat16()
    bias = torch.randn((N,), device="cuda")
    out = torch.zeros((M, N), device="cuda", dtype=torch.float32 if mode == 0 else torch.bfloat16)
    resid = out if inplace else None
    def call():
        check(lib, lib.parseq_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), M, N, K, mode, 1.0,
                                        resid.data_ptr() if resid is not None else None, N if resid is not None else 0, 0, out.data_ptr(), N, st))
    for _ in range(3): call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): call()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1000 / iters
    return us, 2.0 * M * N * K / us / 1e6
print(f"{'shape':44s} |  cg=1 us   TF/s |  cg=2 us   TF/s | cg2/cg1")
for name, M, N, K, mode, ip in [("S qkv 512img", 65536, 1152, 384, 1, False), ("S fc1 512img", 65536, 1536, 384, 2, False),
                                ("S fc2 unfused 256img", 32768, 384, 1536, 0, True), ("S proj unfused 256img", 32768, 384, 384, 0, True),
                                ("B48 qkv 256img (T=240)", 61440, 2304, 768, 1, False), ("B48 proj", 61440, 768, 768, 0, True),
                                ("B48 fc1", 61440, 3072, 768, 2, False), ("B48 fc2", 61440, 768, 3072, 0, True),
                                ("S cross-KV 512img", 65536, 768, 384, 1, False), ("K=4096 N=1536", 16384, 1536, 4096, 1, False)]:
    u1, t1 = run(M, N, K, mode, ip, 1)
    u2, t2 = run(M, N, K, mode, ip, 2)
    print(f"{name:24s} M={M:6d} N={N:5d} K={K:5d} | {u1:8.1f} {t1:7.1f} | {u2:8.1f} {t2:7.1f} | {u2 / u1:6.3f}")
check(lib, lib.parseq_set_option(None, b"cta_group", 0)); check(lib, lib.parseq_set_option(None, b"block_n", 0))
