"""Turns the ncu exports of tests/gpu_r2_profile.sh (gpurun_out/r2_*.csv) into the tracked summaries under profiles/:
launch list with shares, per-kernel key metrics of the --set full captures, and the DRAM-traffic JSON bench.py reads."""
import csv, io, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")

def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("pq::", "")

# ---- launch list ----
lines = [l for l in open(os.path.join(G, "r2_launches_bs512.csv")) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg, order = {}, []
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
total = sum(a[1] for a in agg.values())
with open(os.path.join(OUT, "r2_launches_bs512.txt"), "w") as f:
    f.write("# ncu launch list, one PARSeq-S forward, bs=512, AR + 1 refine (cold-cache, serialised per-launch times: compare SHARES)\n")
    f.write("# cmd: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tests/profile_step.py 512\n")
    f.write(f"{'kernel':54s} {'n':>3s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:54s} {n:3d} {us:10.1f} {us / n:9.2f} {us / total:6.3f}\n")
    f.write(f"total_us {total:.1f}  launches {sum(a[0] for a in agg.values())}\n")
print(open(os.path.join(OUT, "r2_launches_bs512.txt")).read())

# ---- full captures ----
KEYS = [("time_us", "gpu__time_duration.sum"), ("dram_read_bytes", "dram__bytes_read.sum"), ("dram_write_bytes", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("registers", "launch__registers_per_thread"), ("inst_executed", "smsp__inst_executed.sum"),
        ("lts_read_bytes", "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum")]
def load(name, labels):
    raw = open(os.path.join(G, f"r2_raw_{name}.csv")).read()
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rr[0], rr[1], rr[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for row, lab in zip(data, labels):
        d = {"kernel": short(row[col["Kernel Name"]]), "what": lab}
        for key, metric in KEYS:
            if metric not in col or row[col[metric]] == "":
                continue
            v = float(row[col[metric]].replace(",", ""))
            u = units[col[metric]]
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
            d[key] = v * mult
        out.append(d)
    return out
inst = []
inst += load("gemm", ["QKV (M=65536,N=1152,K=384)", "fc1+GELU (N=1536,K=384), FFMA2 epilogue"])
inst += load("gemm_ln", ["attn.proj + residual + norm2 (K=384)", "mlp.fc2 + residual + next norm1 (K=1536)"])
inst += load("ar2", ["whole AR loop: 26 steps, 512 images, 22 clusters of 6"])
inst += load("attn", ["ViT attention core, 512 x 6 (image, head) CTAs"])
if os.path.exists(os.path.join(G, "r2_raw_mlp_ln.csv")):
    inst += load("mlp_ln", ["OPT-IN one-kernel MLP: fc1 + GELU + fc2 + residual + next norm1 (replaces fc1 + fc2 rows)"])
def dram(d): return d.get("dram_read_bytes", 0.0) + d.get("dram_write_bytes", 0.0)
traffic = {
    "source": "ncu --set full --clock-control none, tests/gpu_r2_profile.sh, bs=512 forward, encoder block 0 / AR loop (round 2 final build)",
    "instances": inst,
    "avg_dram_bytes_per_launch": (dram(inst[0]) + dram(inst[1])) / 2,
    "fused_avg_dram_bytes_per_launch": (dram(inst[2]) + dram(inst[3])) / 2,
    "dec_ar_dram_bytes_per_launch": dram(inst[4]),
    "attn_dram_bytes_per_launch": dram(inst[5]),
}
json.dump(traffic, open(os.path.join(OUT, "r2_ncu_traffic.json"), "w"), indent=1)
with open(os.path.join(OUT, "r2_ncu_summary.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none (tests/gpu_r2_profile.sh), PARSeq-S bs=512 forward, round-2 final build\n")
    f.write(f"{'kernel / what':92s} {'us':>8s} {'DRAM MB':>8s} {'dram%':>6s} {'tensor%':>7s} {'sm%':>6s} {'issue%':>6s} {'L2hit%':>6s} {'regs':>5s} {'Minst':>7s}\n")
    for d in inst:
        f.write(f"{(d['kernel'] + ' | ' + d['what'])[:92]:92s} {d.get('time_us', 0):8.1f} {dram(d) / 1e6:8.1f} {d.get('dram_pct', 0):6.1f} "
                f"{d.get('tensor_pipe_pct', 0):7.1f} {d.get('sm_pct', 0):6.1f} {d.get('issue_active_pct', 0):6.1f} {d.get('l2_hit_pct', 0):6.1f} "
                f"{int(d.get('registers', 0)):5d} {d.get('inst_executed', 0) / 1e6:7.1f}\n")
print(open(os.path.join(OUT, "r2_ncu_summary.txt")).read())
