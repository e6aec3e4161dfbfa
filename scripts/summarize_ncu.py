"""Summarise ncu artefacts from gpurun_out/ into profiles/ (small, tracked text/JSON).
   python scripts/summarize_ncu.py <tag>      (reads gpurun_out/launches_<tag>.csv, prof_{gemm,attn,misc}_<tag>.ncu-rep)"""
import collections, csv, io, json, os, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def launches():
    lines = [l for l in open(os.path.join(G, f"launches_{tag}.csv")) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] in ("usecond", "us") else (v / 1e6 if row["Metric Unit"] in ("nsecond", "ns") else v)
        k = (row["Kernel Name"].split("(")[0][:70], row["Grid Size"], row["Block Size"])
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    out = [f"# ncu launch list of ONE bench step (batch 64 x 30 s): gpu__time_duration per launch, cold-cache and serialised\n"
           f"# command: scripts/gpu_ncu.sh {tag} (ncu --metrics gpu__time_duration.sum --clock-control none -k <our kernels>, the 4th step of python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-alt-precision --no-config5)\n"
           f"# total {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches\n"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{a[1]:9.3f} ms {100*a[1]/tot:5.1f}%  {a[0]:4d}x  avg {a[1]/a[0]:8.4f} ms  {k[0]}  grid={k[1]} block={k[2]}")
    open(os.path.join(P, f"{tag}_launches.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:12]))

KEYS = ["smsp__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_tensor.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

def raw(name):
    rep = os.path.join(G, f"prof_{name}_{tag}.ncu-rep")
    csvp = os.path.join(G, f"raw_{name}_{tag}.csv")
    if os.path.exists(csvp) and os.path.getsize(csvp) > 0:
        txt = open(csvp).read()
    elif os.path.exists(rep):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return []
    r = list(csv.reader(io.StringIO(txt)))
    hdr, units, rows = r[0], r[1], r[2:]
    res = []
    for row in rows:
        d = {"kernel": row[hdr.index("Kernel Name")].split("(")[0]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(row[i].replace(",", ""))
                except ValueError:
                    continue
                if units[i] in UNIT:
                    v *= UNIT[units[i]]
                d[k + (" [" + units[i] + "]" if units[i] not in UNIT and units[i] else "")] = v
        res.append(d)
    return res

launches()
summary = {n: raw(n) for n in ("gemm", "attn", "attn_ragged", "rvq_search", "rvq_sgemm", "layernorm", "attn_mma", "misc")}
summary = {n: v for n, v in summary.items() if v}
json.dump(summary, open(os.path.join(P, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
# DRAM traffic per launch of the dominant kernel class (mean over the captured layer's four GEMM launches)
g = [d for d in summary.get("gemm", []) if "dram__bytes_read.sum" in d]
traffic = {}
if g:
    traffic["gemm_bf16_tcgen05"] = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in g) / len(g)
a = [d for d in summary.get("attn", []) if "dram__bytes_read.sum" in d]
if a:
    traffic["attention_encoder"] = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in a) / len(a)
traffic["_source"] = f"profiles/{tag}_ncu_full_summary.json (ncu --set full, mean per launch over the captured launches)"
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
for n, rows in summary.items():
    for d in rows:
        print(n, d["kernel"][-40:], {k.split(".")[0][-28:]: (round(v, 3) if v < 1e4 else f"{v:.3e}") for k, v in d.items() if k != "kernel"})
