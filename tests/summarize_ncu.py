"""Turns the ncu outputs of tests/gpu_profile2.sh (gpurun_out/) into the tracked summaries under profiles/:
launch list with shares, per-kernel key metrics of the --set full capture, and the DRAM-traffic JSON bench.py reads."""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "v5"
out_dir = os.path.join(ROOT, "profiles")

def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("pq::", "")

# ---- launch list ----
path = os.path.join(ROOT, "gpurun_out", f"launches_bs512_{tag}.csv")
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = {}
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
with open(os.path.join(out_dir, f"r1_launches_bs512_{tag}.txt"), "w") as f:
    f.write("# ncu launch list, one PARSeq-S forward, bs=512, AR + 1 refine (cold-cache, serialised per-launch times: compare SHARES)\n")
    f.write("# cmd: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tests/profile_step.py 512\n")
    f.write(f"{'kernel':54s} {'n':>3s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:54s} {n:3d} {us:10.1f} {us / n:9.2f} {us / total:6.3f}\n")
    f.write(f"total_us {total:.1f}  launches {sum(a[0] for a in agg.values())}\n")
print(open(os.path.join(out_dir, f"r1_launches_bs512_{tag}.txt")).read())

# ---- full capture ----
rep = os.path.join(ROOT, "gpurun_out", f"prof_gemm_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rr[0], rr[1], rr[2:]
col = {h: i for i, h in enumerate(hdr)}
def val(row, name):
    i = col[name]
    v = float(row[i].replace(",", ""))
    u = units[i]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
    return v * mult
labels = ["QKV (M=65536,N=1152,K=384)", "attn.proj + residual + norm2 (K=384)", "fc1+GELU (N=1536,K=384)",
          "mlp.fc2 + residual + next norm1 (K=1536)"]
inst = []
for row, lab in zip(data, labels):
    inst.append({
        "kernel": short(row[col["Kernel Name"]]), "gemm": lab,
        "time_us": round(val(row, "gpu__time_duration.sum"), 2),
        "dram_read_bytes": val(row, "dram__bytes_read.sum"), "dram_write_bytes": val(row, "dram__bytes_write.sum"),
        "tensor_pipe_pct": float(row[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
        "dram_pct": float(row[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
        "sm_pct": float(row[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]]),
        "issue_active_pct": float(row[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
        "registers": int(float(row[col["launch__registers_per_thread"]])),
        "inst_executed": float(row[col["smsp__inst_executed.sum"]].replace(",", "")),
    })
plain = [i for i in inst if "gemm_bf16" in i["kernel"]]
fused = [i for i in inst if "gemm_ln" in i["kernel"]]
avg = lambda xs: sum(i["dram_read_bytes"] + i["dram_write_bytes"] for i in xs) / max(1, len(xs))
blob = {"source": f"ncu --set full --clock-control none, tests/gpu_profile2.sh, bs=512 forward, encoder block 0 ({tag})",
        "instances": inst, "avg_dram_bytes_per_launch": avg(plain), "fused_avg_dram_bytes_per_launch": avg(fused)}
json.dump(blob, open(os.path.join(out_dir, "r1_gemm_ncu_traffic.json"), "w"), indent=1)
with open(os.path.join(out_dir, f"r1_ncu_summary_{tag}.txt"), "w") as f:
    f.write("# ncu --set full of the four encoder GEMM kinds of block 0 in one bs=512 forward (cmd lines in tests/gpu_profile2.sh)\n")
    for i in inst:
        f.write("  " + " | ".join(f"{k}={v}" for k, v in i.items()) + "\n")
print(open(os.path.join(out_dir, f"r1_ncu_summary_{tag}.txt")).read())
