"""Microbenchmark (not a test): the one-kernel MLP + LayerNorm (mlp_ln.cuh) vs fc1 (GELU epilogue) + fused fc2/LayerNorm."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
lib = load_library()
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / iters

def run(M, D):
    H = 4 * D
    xn = torch.randn((M, D), device="cuda").bfloat16()
    W1 = (torch.randn((H, D), device="cuda") * 0.05).bfloat16(); b1 = torch.randn((H,), device="cuda")
    W2 = (torch.randn((D, H), device="cuda") * 0.03).bfloat16(); b2 = torch.randn((D,), device="cuda")
    g = torch.ones((D,), device="cuda"); b = torch.zeros((D,), device="cuda")
    x = torch.randn((M, D), device="cuda"); xo = torch.empty((M, D), device="cuda", dtype=torch.bfloat16)
    hid = torch.empty((M, H), device="cuda", dtype=torch.bfloat16)
    fused = lambda: check(lib, lib.parseq_mlp_ln_bf16(xn.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, D,
                                                      x.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-6, xo.data_ptr(), st))
    fc1 = lambda: check(lib, lib.parseq_gemm_bf16(xn.data_ptr(), D, W1.data_ptr(), D, b1.data_ptr(), M, H, D, 2, 1.0, None, 0, 0,
                                                  hid.data_ptr(), H, st))
    fc2 = lambda: check(lib, lib.parseq_gemm_ln_bf16(hid.data_ptr(), H, W2.data_ptr(), H, b2.data_ptr(), M, D, H, x.data_ptr(),
                                                     g.data_ptr(), b.data_ptr(), 1e-6, xo.data_ptr(), st))
    tf, t1, t2 = timeit(fused), timeit(fc1), timeit(fc2)
    fl = 4.0 * M * D * H
    print(f"M={M:6d} D={D} | one kernel {tf:7.1f} us ({fl/tf/1e6:6.1f} TF/s) | fc1 {t1:6.1f} + fc2/LN {t2:6.1f} = {t1+t2:7.1f} us "
          f"({fl/(t1+t2)/1e6:6.1f} TF/s)", flush=True)

for cg, pdl in ((1, 0), (2, 0), (2, 1)):
    check(lib, lib.parseq_set_option(None, b"mlp_cta_group", cg))
    check(lib, lib.parseq_set_option(None, b"pair_pdl", pdl))
    print(f"--- one kernel: mlp_cta_group={cg} pair_pdl={pdl}", flush=True)
    for M in (148 * 128, 65536, 148 * 128 * 4):
        run(M, 384)
    run(65536, 192)
