#!/usr/bin/env python
"""Benchmark of the B200-native PARSeq engine: images/sec, PARSeq-S 32x128, AR + 1 refine iteration
(BASELINE.json metric; config[1]: bs=512 per GPU, bf16 tensor-core operands, synthetic crops / seeded
random weights), p50 latency at bs=1, tensor-roofline fraction of the tcgen05 GEMM kernel and the
reference-style CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  "step" = one forward of `batch` images per GPU (weak scaling:
images are independent, no data-path collective; model.py:105-169 has no cross-image dependency).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_GFLOP_PER_IMAGE = 6.038   # PARSeq-S, AR + 1 refine, 2 FLOP per MAC


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="images per GPU per step")
    ap.add_argument("--chunk", type=int, default=0, help="images per pipeline stage (0 = engine default)")
    ap.add_argument("--max-batch", type=int, default=0, help="images per super-chunk / CUDA graph (0 = default)")
    ap.add_argument("--dec-chunk", type=int, default=0, help="images per decoder chain (0 = engine default 128)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--fuse-ln", type=int, default=-1, help="bit 0: attn.proj, bit 1: mlp.fc2 fused with the following LayerNorm (engine default 3)")
    ap.add_argument("--attn-impl", type=int, default=-1, help="encoder attention: 1 tcgen05 (default), 0 mma.sync")
    ap.add_argument("--no-ar-kernel", action="store_true", help="AR loop as separate kernels instead of the persistent kernel")
    ap.add_argument("--cta-group", type=int, default=0, help="GEMM tile: 0 auto, 1 single CTA, 2 CTA pair")
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = f"/tmp/parseq_clocks_{os.getpid()}.csv"

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def pick_cpu_threads(o, cfg):
    """torch's intra-op pool is not automatically fastest at os.cpu_count() threads on a many-core host (the
    decoder's small matmuls oversubscribe); give the CPU arm its best thread count from a short probe."""
    import torch
    from parseq_b200.weights import synth_images
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    probe = synth_images(cfg, 16, 7)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        o.forward(probe[:4], None, True, 1)
        t0 = time.perf_counter(); o.forward(probe, None, True, 1); dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def oracle_images_per_sec(cfg, sd, batch, repeats, decode_ar=True, refine_iters=1):
    """Reference-style CPU path: fp32 restatement of model.py:105-169 (oracle, pinned to the reference's
    own modules by tests/golden) on the host cores (best thread count of a short probe)."""
    import torch
    from oracle.parseq_oracle import ParseqOracle
    from parseq_b200.weights import synth_images
    o = ParseqOracle(cfg, sd, "fp32")
    pick_cpu_threads(o, cfg)
    x = synth_images(cfg, batch, 4242)
    o.forward(x[: max(1, batch // 8)], None, decode_ar, refine_iters)     # warm-up
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.forward(x, None, decode_ar, refine_iters)
        ts.append(time.perf_counter() - t0)
    return batch / statistics.median(ts), torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; /root/reference
    does not exist on the GPU box) on the host cores, same metric/config, bounded sample per step."""
    import torch
    from parseq_b200.config import make_config
    from parseq_b200.weights import init_state_dict, synth_images
    from oracle.parseq_oracle import ParseqOracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_config("parseq")
    sd = init_state_dict(cfg, 0)
    o = ParseqOracle(cfg, sd, "fp32")
    pick_cpu_threads(o, cfg)
    probe = synth_images(cfg, 8, 1)
    o.forward(probe, None, True, 1)
    t0 = time.perf_counter(); o.forward(probe, None, True, 1); dt = time.perf_counter() - t0
    ips0 = 8 / dt
    budget_s = 150.0
    sample = int(max(1, min(args.batch, ips0 * budget_s / max(1, args.steps + args.warmup))))
    x = synth_images(cfg, sample, 2)
    for _ in range(args.warmup):
        o.forward(x, None, True, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.forward(x, None, True, 1)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "images/sec PARSeq-S 32x128 AR+1refine", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PARSeq-S 32x128 bs=512 AR+1refine (configs[1]); CPU arm steps over a bounded "
                               f"sample of {sample} images", "batch_per_step": sample},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images/step x {args.steps} steps, fp32 torch CPU oracle (pinned to "
                                   "reference modules via tests/golden)"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from parseq_b200.config import make_config
    from parseq_b200.factory import create_model
    from parseq_b200.weights import init_state_dict, synth_images

    cfg = make_config("parseq")
    sd = init_state_dict(cfg, 0)
    model = create_model("parseq", decode_ar=True, refine_iters=1)
    model.model.load_state_dict(sd)
    if args.max_batch:
        model.model.set_engine_option("max_batch", args.max_batch)
    if args.chunk:
        model.model.set_engine_option("chunk", args.chunk)
    if args.dec_chunk:
        model.model.set_engine_option("dec_chunk", args.dec_chunk)
    if args.no_graph:
        model.model.set_engine_option("use_graph", 0)
    if args.no_pdl:
        model.model.set_engine_option("pdl", 0)
    if args.attn_impl >= 0:
        model.model.set_engine_option("attn_impl", args.attn_impl)
    if args.no_ar_kernel:
        model.model.set_engine_option("ar_kernel", 0)
    if args.fuse_ln >= 0:
        model.model.set_engine_option("fuse_ln", args.fuse_ln)
    if args.cta_group:
        model.model.set_engine_option("cta_group", args.cta_group)
    if args.block_n:
        model.model.set_engine_option("block_n", args.block_n)
    model = model.eval().to(dev)
    eng = model.model.engine()
    B = args.batch
    st = torch.cuda.current_stream(dev)

    # Inputs larger than L2: NROT distinct resident batches (NROT * 25.2 MB > 126 MB), rotated per step.
    NROT = 8
    batches = [synth_images(cfg, B, 100 + rank * NROT + i).to(dev) for i in range(NROT)]
    logits = torch.empty((B, 26, cfg.num_classes), dtype=torch.float32, device=dev)
    ids = torch.empty((B, 26), dtype=torch.int32, device=dev)
    steps_t = torch.empty((1,), dtype=torch.int32, device=dev)

    def step(i):
        eng.forward(batches[i % NROT].data_ptr(), B, logits.data_ptr(), ids.data_ptr(), steps_t.data_ptr(),
                    st.cuda_stream, None, True, 1)

    for i in range(max(3, args.warmup)):
        step(i)
    torch.cuda.synchronize(dev)
    if distributed:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(st)
    for i in range(args.steps):
        step(i)
    e1.record(st)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    launches = eng.launches - l0
    if distributed:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else {}
    value = world * B * args.steps / (ms / 1000.0)

    # ---- end-to-end through the host-buffer entry point (pinned host memory, H2D + D2H inside) ----
    himg = [synth_images(cfg, B, 500 + rank * 2 + i).pin_memory() for i in range(2)]
    hlog = torch.empty((B, 26, cfg.num_classes), dtype=torch.float32).pin_memory()
    hids = torch.empty((B, 26), dtype=torch.int32).pin_memory()
    hsteps = torch.empty((1,), dtype=torch.int32).pin_memory()
    for i in range(2):
        eng.forward_host(himg[i % 2].data_ptr(), B, hlog.data_ptr(), hids.data_ptr(), hsteps.data_ptr(),
                         st.cuda_stream, None, True, 1)
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.forward_host(himg[i % 2].data_ptr(), B, hlog.data_ptr(), hids.data_ptr(), hsteps.data_ptr(),
                         st.cuda_stream, None, True, 1)
    e2e_s = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * B * args.steps / e2e_s
    # extra: same end-to-end call with raw uint8 HWC crops (transform folded into the patch gather)
    hu8 = [torch.randint(0, 256, (B, cfg.img_size[0], cfg.img_size[1], 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for i in range(2):
        eng.forward_u8(hu8[i % 2].data_ptr(), B, hlog.data_ptr(), hids.data_ptr(), hsteps.data_ptr(), st.cuda_stream, None,
                       True, 1, host=True)
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.forward_u8(hu8[i % 2].data_ptr(), B, hlog.data_ptr(), hids.data_ptr(), hsteps.data_ptr(), st.cuda_stream, None,
                       True, 1, host=True)
    e2e_u8_s = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([e2e_u8_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_u8_s = float(t.item())
    e2e_u8_val = world * B * args.steps / e2e_u8_s
    h2d = B * 3 * cfg.img_size[0] * cfg.img_size[1] * 4
    d2h = B * 26 * cfg.num_classes * 4 + B * 26 * 4 + 4

    if distributed:
        dist.barrier()
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---- roofline pass: per-category device time from CUDA-event pairs around every launch ----
    peaks, peak_src = load_peaks()
    eng.set_option("timing", 1)
    step(0)
    torch.cuda.synchronize(dev)
    tim = eng.get_timing()
    eng.set_option("timing", 0)
    enc = tim["enc_gemm"]
    gemm_tflops = enc["flops"] / (enc["ms"] * 1e-3) / 1e12 if enc["ms"] > 0 else 0.0
    # The GEMM kernel is timed inside a long step -> sustained cuBLAS figure is the denominator
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    total_timed = sum(v["ms"] for v in tim.values())
    ncu = {}
    tpath = os.path.join(ROOT, "profiles", "r1_gemm_ncu_traffic.json")
    if os.path.exists(tpath):          # dram__bytes_read+write per launch from the committed ncu --set full capture
        with open(tpath) as f:
            ncu = json.load(f)
    by_cat = {k: round(v["ms"], 4) for k, v in tim.items()}
    whole = (value / world) * ALG_GFLOP_PER_IMAGE * 1e9 / (peak_tf * 1e12)
    rl_gemm = {
        "bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (QKV, fc1+GELU, patch embedding, cross K/V)",
        "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tflops / peak_tf,
        "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": ncu.get("avg_dram_bytes_per_launch"),
        "traffic_source": "profiles/r1_gemm_ncu_traffic.json (ncu --set full, avg dram bytes per launch of QKV / fc1)",
        "flops_per_launch": enc["flops"] / max(1, enc["launches"]),
        "avg_launch_ms": enc["ms"] / max(1, enc["launches"]),
        "share_of_step": enc["ms"] / total_timed if total_timed else None,
    }
    fus = tim.get("enc_gemm_ln", {"ms": 0.0, "flops": 0.0, "launches": 0})
    rl_fused = None
    if fus["launches"] > 0 and fus["ms"] > 0:
        # algorithmic bytes of x += A W^T + b ; xn = LN(x): A (bf16) + W (bf16) + x read and written (fp32) + xn (bf16);
        # per encoder block one launch with K = D (attn.proj) and one with K = 4 D (mlp.fc2)
        Mrows, D_ = args.batch * cfg.enc_tokens, cfg.embed_dim
        def fused_bytes(K):
            return Mrows * K * 2 + D_ * K * 2 + 2 * Mrows * D_ * 4 + Mrows * D_ * 2
        per_block = fused_bytes(D_) + fused_bytes(D_ * cfg.enc_mlp_ratio)
        alg_bytes = per_block * (fus["launches"] / 2.0)
        gbs = alg_bytes / (fus["ms"] * 1e-3) / 1e9
        rl_fused = {
            "bound": "hbm", "kernel": "gemm_ln_fused_kernel (attn.proj / mlp.fc2 + residual + following LayerNorm)",
            "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
            "peak_source": f"{peak_src} hbm_gbs", "traffic": ncu.get("fused_avg_dram_bytes_per_launch"),
            "traffic_source": "profiles/r1_gemm_ncu_traffic.json (ncu --set full, avg dram bytes per launch of the two fused GEMMs)",
            "bytes_per_launch": alg_bytes / fus["launches"], "avg_launch_ms": fus["ms"] / fus["launches"],
            "tflops": fus["flops"] / (fus["ms"] * 1e-3) / 1e12,
            "share_of_step": fus["ms"] / total_timed if total_timed else None,
        }
    # the dominant kernel (largest share of the step) is the headline roofline; the other one rides along
    if rl_fused is not None and rl_fused["share_of_step"] > rl_gemm["share_of_step"]:
        roofline, other = rl_fused, rl_gemm
    else:
        roofline, other = rl_gemm, rl_fused
    roofline = dict(roofline)
    roofline["by_category_ms"] = by_cat
    roofline["whole_step_frac_of_tensor_peak"] = whole
    roofline["other_kernel"] = other

    # ---- p50 latency at bs=1 ----
    lat = None
    if not args.no_latency and world == 1:
        x1 = synth_images(cfg, 1, 9).to(dev)
        l1 = torch.empty((1, 26, cfg.num_classes), dtype=torch.float32, device=dev)
        i1 = torch.empty((1, 26), dtype=torch.int32, device=dev)
        for _ in range(10):
            eng.forward(x1.data_ptr(), 1, l1.data_ptr(), i1.data_ptr(), steps_t.data_ptr(), st.cuda_stream, None, True, 1)
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(200):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            eng.forward(x1.data_ptr(), 1, l1.data_ptr(), i1.data_ptr(), steps_t.data_ptr(), st.cuda_stream, None, True, 1)
            b.record(st)
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        lat = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[int(len(ts) * 0.99) - 1], "iters": len(ts)}

    # ---- CPU baseline on this box's host cores (bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sample = 96
        ips, cores = oracle_images_per_sec(cfg, sd, sample, 2)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"median of 2 forwards of {sample} images (of the 512-image workload), fp32 torch CPU "
                         "oracle of model.py:105-169"}

    line = {
        "metric": "images/sec PARSeq-S 32x128 AR+1refine", "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "PARSeq-S 32x128 94-char max_len=25 bs=512/GPU AR + 1 refine (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world} (batch-sharded, no collective)",
                   "l2": f"inputs rotate over {NROT} resident batches ({NROT * B * 49152 / 1e6:.0f} MB > 126 MB L2)",
                   "chunk": args.chunk or (args.max_batch or 512), "max_batch": args.max_batch or 512, "dec_chunk": args.dec_chunk or 128, "cuda_graph": not args.no_graph},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1000 * e2e_s / args.steps},
        "e2e_u8": {"value": e2e_u8_val, "unit": "images/s", "h2d_bytes_per_step": h2d // 4, "d2h_bytes_per_step": d2h,
                   "note": "parseq_forward_host_u8: raw uint8 HWC crops, ToTensor+Normalize folded into the patch gather"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "latency_bs1": lat,
    }
    print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
