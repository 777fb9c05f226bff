"""Epilogue variants of the K=384 GEMMs: TMA-store epilogue vs direct stores, single CTA vs pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
lib = load_library()
st = torch.cuda.current_stream().cuda_stream
def run(M, N, K, mode, cg, tma, bn=0, stages=0, iters=20):
    for k, v in ((b"cta_group", cg), (b"block_n", bn), (b"tma_epilogue", tma), (b"gemm_stages", stages)):
        check(lib, lib.parseq_set_option(None, k, v))
    A = torch.randn((M, K), device="cuda").bfloat16()
    W = (torch.randn((N, K), device="cuda") * 0.02).bfloat16()
    bias = torch.randn((N,), device="cuda")
    out = torch.zeros((M, N), device="cuda", dtype=torch.float32 if mode == 0 else torch.bfloat16)
    def call():
        check(lib, lib.parseq_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), M, N, K, mode, 1.0, None, 0, 0, out.data_ptr(), N, st))
    for _ in range(3): call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): call()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / iters
for name, N, mode in (("qkv", 1152, 1), ("fc1+gelu", 1536, 2), ("fc1 no act", 1536, 1)):
    for cg in (1, 2):
        for bn in (256, 128):
            for tma in (1, 0):
                us = run(65536, N, 384, mode, cg, tma, bn)
                print(f"{name:10s} cg={cg} bn={bn} {'tma-store' if tma else 'direct   '} : {us:7.1f} us  {2.0*65536*N*384/us/1e6:7.1f} TF/s")
for k, v in ((b"cta_group", 0), (b"block_n", 0), (b"tma_epilogue", 1), (b"gemm_stages", 0)):
    check(lib, lib.parseq_set_option(None, k, v))
