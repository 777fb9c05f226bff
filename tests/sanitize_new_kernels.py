"""compute-sanitizer driver for the kernels added late in round 2: the one-kernel MLP + LayerNorm (single CTA and CTA pair)
and the CTA-pair variant of the fused GEMM + LayerNorm, at ragged sizes (several tiles per CTA pair, partial last tile),
checked against the two-kernel path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
lib = load_library()
st = torch.cuda.current_stream().cuda_stream

for D, M in ((384, 300), (192, 513), (384, 148 * 128 + 77)):
    H = 4 * D
    g = torch.Generator(device="cuda").manual_seed(M)
    xn = torch.randn((M, D), device="cuda", generator=g).bfloat16()
    W1 = (torch.randn((H, D), device="cuda", generator=g) * 0.06).bfloat16(); b1 = torch.randn((H,), device="cuda", generator=g)
    W2 = (torch.randn((D, H), device="cuda", generator=g) * 0.04).bfloat16(); b2 = torch.randn((D,), device="cuda", generator=g)
    ga = torch.ones((D,), device="cuda"); be = torch.zeros((D,), device="cuda")
    x0 = torch.randn((M, D), device="cuda", generator=g)
    hid = torch.empty((M, H), dtype=torch.bfloat16, device="cuda")
    check(lib, lib.parseq_gemm_bf16(xn.data_ptr(), D, W1.data_ptr(), D, b1.data_ptr(), M, H, D, 2, 1.0, None, 0, 0, hid.data_ptr(), H, st))
    ref = {}
    check(lib, lib.parseq_set_option(None, b"ln_split", 1))
    for cg in (1, 2):
        check(lib, lib.parseq_set_option(None, b"ln_cta_group", cg))
        x = x0.clone(); xo = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
        check(lib, lib.parseq_gemm_ln_bf16(hid.data_ptr(), H, W2.data_ptr(), H, b2.data_ptr(), M, D, H, x.data_ptr(), ga.data_ptr(),
                                           be.data_ptr(), 1e-6, xo.data_ptr(), st))
        torch.cuda.synchronize()
        ref[cg] = (x, xo)
    assert torch.equal(ref[1][0], ref[2][0]) and torch.equal(ref[1][1], ref[2][1])
    if D == 384:                                       # column-split pair kernel (gemm_ln2.cuh): x identical, xn within one bf16 ulp
        check(lib, lib.parseq_set_option(None, b"ln_split", 2))
        x = x0.clone(); xo = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
        check(lib, lib.parseq_gemm_ln_bf16(hid.data_ptr(), H, W2.data_ptr(), H, b2.data_ptr(), M, D, H, x.data_ptr(), ga.data_ptr(),
                                           be.data_ptr(), 1e-6, xo.data_ptr(), st))
        torch.cuda.synchronize()
        assert torch.equal(x, ref[1][0])
        assert (xo.float() - ref[1][1].float()).abs().max().item() <= 2.0 ** -7 * ref[1][1].float().abs().max().item()
    check(lib, lib.parseq_set_option(None, b"ln_split", 1))
    for cg in (1, 2):
        check(lib, lib.parseq_set_option(None, b"mlp_cta_group", cg))
        x = x0.clone(); xo = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
        check(lib, lib.parseq_mlp_ln_bf16(xn.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, D, x.data_ptr(),
                                          ga.data_ptr(), be.data_ptr(), 1e-6, xo.data_ptr(), st))
        torch.cuda.synchronize()
        assert torch.equal(x, ref[1][0]) and torch.equal(xo, ref[1][1])
    print("ok: D", D, "M", M, "gemm_ln pair == single, mlp_ln (single, pair) == two kernels", flush=True)
check(lib, lib.parseq_set_option(None, b"ln_split", 0))
print("sanitize_new_kernels done")
