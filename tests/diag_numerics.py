"""Diagnostic (not a test): where does engine-vs-oracle error come from?  Prints only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parseq_b200.build import build
build()
from parseq_b200.engine import load_library, check
from parseq_b200.config import make_config
from parseq_b200.factory import create_model
from parseq_b200.weights import init_state_dict, synth_images
from oracle.parseq_oracle import ParseqOracle

lib = load_library()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
for (M, N, K) in [(1024, 1152, 384), (1024, 384, 1536), (1024, 384, 96)]:
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    W = (torch.randn((N, K), device="cuda", generator=g) * 0.02).bfloat16()
    out = torch.empty((M, N), device="cuda")
    check(lib, lib.parseq_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, None, M, N, K, 0, 1.0, None, 0, 0, out.data_ptr(), N, st))
    torch.cuda.synchronize()
    ref64 = A.double() @ W.double().t()
    ref32 = (A.float() @ W.float().t()).double()
    scale = ref64.abs().mean().item()
    print(f"GEMM {M}x{N}x{K}: engine-vs-fp64 max {((out.double()-ref64).abs().max()/scale).item():.2e} "
          f"mean {((out.double()-ref64).abs().mean()/scale).item():.2e} | torch-fp32-vs-fp64 max "
          f"{((ref32-ref64).abs().max()/scale).item():.2e} mean {((ref32-ref64).abs().mean()/scale).item():.2e}")

for depth in (0, 1, 2, 4, 12):
    cfg = make_config("parseq", enc_depth=depth) if depth else make_config("parseq", enc_depth=1)
    sd = init_state_dict(cfg, 0)
    m = create_model("parseq", enc_depth=cfg.enc_depth)
    m.model.load_state_dict(sd)
    m = m.eval().to("cuda")
    x = synth_images(cfg, 2, 0)
    with torch.inference_mode():
        mem = m.model.encode(x.cuda()).cpu()
    ob = ParseqOracle(cfg, sd, "bf16"); o32 = ParseqOracle(cfg, sd, "fp32")
    mb, blocks_b = ob.encode(x, return_blocks=True)
    m32 = o32.encode(x)
    print(f"enc_depth={cfg.enc_depth}: |engine-bf16oracle| max {(mem-mb).abs().max():.3e} mean {(mem-mb).abs().mean():.3e}"
          f" | |engine-fp32| max {(mem-m32).abs().max():.3e} mean {(mem-m32).abs().mean():.3e}"
          f" | |bf16oracle-fp32| max {(mb-m32).abs().max():.3e} mean {(mb-m32).abs().mean():.3e}  (|mem| mean {m32.abs().mean():.2f})")

cfg = make_config("parseq"); sd = init_state_dict(cfg, 0)
for ar, ri in [(False, 0), (True, 0), (True, 1)]:
    m = create_model("parseq", decode_ar=ar, refine_iters=ri); m.model.load_state_dict(sd); m = m.eval().to("cuda")
    x = synth_images(cfg, 8, 0)
    with torch.inference_mode():
        lg = m.model.forward(m.tokenizer, x.cuda(), 25).cpu()
    ob = ParseqOracle(cfg, sd, "bf16").forward(x, 25, ar, ri)
    o32 = ParseqOracle(cfg, sd, "fp32").forward(x, 25, ar, ri)
    print(f"ar={ar} refine={ri}: |eng-bf16o| max {(lg-ob.logits).abs().max():.3e} mean {(lg-ob.logits).abs().mean():.3e} "
          f"| |eng-fp32| max {(lg-o32.logits).abs().max():.3e} mean {(lg-o32.logits).abs().mean():.3e} "
          f"| |bf16o-fp32| max {(ob.logits-o32.logits).abs().max():.3e} mean {(ob.logits-o32.logits).abs().mean():.3e} "
          f"ids eq bf16o {(lg.argmax(-1)==ob.ids).float().mean():.4f} fp32 {(lg.argmax(-1)==o32.ids).float().mean():.4f}")
