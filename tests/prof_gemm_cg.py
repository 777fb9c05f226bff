"""ncu driver: one QKV-shaped and one K=4096 GEMM with the single-CTA and the CTA-pair tile (cta_group 1 / 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
lib = load_library()
st = torch.cuda.current_stream().cuda_stream
def run(M, N, K, cg):
    check(lib, lib.parseq_set_option(None, b"cta_group", cg))
    check(lib, lib.parseq_set_option(None, b"block_n", 256))
    A = torch.randn((M, K), device="cuda").bfloat16()
    W = (torch.randn((N, K), device="cuda") * 0.02).bfloat16()
    bias = torch.randn((N,), device="cuda")
    out = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        check(lib, lib.parseq_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), M, N, K, 1, 1.0, None, 0, 0, out.data_ptr(), N, st))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    check(lib, lib.parseq_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), M, N, K, 1, 1.0, None, 0, 0, out.data_ptr(), N, st))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
for M, N, K in ((65536, 1152, 384), (16384, 1536, 4096)):
    for cg in (1, 2):
        run(M, N, K, cg)
print("done")
