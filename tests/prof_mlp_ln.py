"""Where the one-kernel MLP (mlp_ln.cuh) waits: cycle counters of CTA 0's MMA thread and of one epilogue warp."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
lib = load_library()
st = torch.cuda.current_stream().cuda_stream

def run(M, D, cg):
    check(lib, lib.parseq_set_option(None, b"mlp_cta_group", cg))
    H = 4 * D
    xn = torch.randn((M, D), device="cuda").bfloat16()
    W1 = (torch.randn((H, D), device="cuda") * 0.05).bfloat16(); b1 = torch.randn((H,), device="cuda")
    W2 = (torch.randn((D, H), device="cuda") * 0.03).bfloat16(); b2 = torch.randn((D,), device="cuda")
    g = torch.ones((D,), device="cuda"); b = torch.zeros((D,), device="cuda")
    x = torch.randn((M, D), device="cuda"); xo = torch.empty((M, D), device="cuda", dtype=torch.bfloat16)
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        check(lib, lib.parseq_mlp_ln_bf16_prof(xn.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, D,
                                               x.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-6, xo.data_ptr(), prof.data_ptr(), st))
    torch.cuda.synchronize()
    pr = prof.cpu().tolist()
    tiles = (M + 128 * cg - 1) // (128 * cg)
    per_cta = (tiles + (148 // cg) - 1) // (148 // cg)
    nc = per_cta * (H // 64)
    us = lambda c: c / 1.9e3          # cycles -> us at ~1.9 GHz (clock64 counts SM clocks)
    print(f"M={M} D={D} cta_group={cg}: {per_cta} tiles per CTA, {nc} chunks")
    print(f"  MMA thread: total {us(pr[0]):7.1f} us | waits: xn tile {us(pr[1]):6.1f}  W1 ring {us(pr[2]):6.1f}  h_full {us(pr[3]):6.1f}  "
          f"peer h_full {us(pr[4]):6.1f}  W2 ring {us(pr[5]):6.1f}  acc2 drained {us(pr[6]):6.1f}")
    print(f"  epilogue warp 0: wait a1_full {us(pr[8]):6.1f}  tmem ld {us(pr[9]):6.1f}  gelu+store+fence+arrive {us(pr[10]):6.1f}  "
          f"wait h_empty {us(pr[11]):6.1f}  LN epilogue {us(pr[12]):6.1f} us   (per chunk: a1 {us(pr[8])/nc:.2f} ld {us(pr[9])/nc:.2f} "
          f"rest {us(pr[10])/nc:.2f} he {us(pr[11])/nc:.2f})", flush=True)

for cg in (1, 2):
    run(148 * 128, 384, cg)
    run(65536, 384, cg)
    run(65536, 192, cg)
