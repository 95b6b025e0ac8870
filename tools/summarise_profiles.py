"""Developer tool: turn the ncu outputs of tools/profile_round.sh into the small, committed summaries under profiles/.
    python tools/summarise_profiles.py r01"""
import csv, io, json, os, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct"]

def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                d[k] = r[hdr.index(k)]
                d.setdefault("units", {})[k] = units[hdr.index(k)]
        out.append(d)
    return out

# launch list -> per-kernel count / mean / share of the step
rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"), errors="replace")) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    per.setdefault(name, []).append(float(r[-1]))
with open(os.path.join(P, f"{tag}_launches_ncu.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ (tools/profile_round.sh); "
            "cold-cache, serialised launches of `bench.py --steps 12 --warmup 3 --no-graph --lanes 1`: compare shares\n")
    f.write("kernel,launches,mean_ns,min_ns,max_ns\n")
    for k, v in per.items():
        f.write(f"{k},{len(v)},{sum(v)/len(v):.0f},{min(v):.0f},{max(v):.0f}\n")
step = {k: sum(v) / len(v) for k, v in per.items() if k not in ("k_anchors", "k_anchor_terms")}
tot = sum(step.values())
summary = {"launch_list_share_of_step": {k: round(v / tot, 4) for k, v in step.items()}, "step_sum_ns": round(tot)}
full = raw(os.path.join(G, f"{tag}_full.ncu-rep"))
cf = raw(os.path.join(G, f"{tag}_cf.ncu-rep")) if os.path.exists(os.path.join(G, f"{tag}_cf.ncu-rep")) else []
summary["ncu_full"] = full + cf
json.dump(summary, open(os.path.join(P, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
for d in full:
    if "k_dense_decode" in d["Kernel Name"]:
        mb = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[d["units"]["dram__bytes_read.sum"]]
        json.dump({"workload": "cfg1", "k_dense_decode_dram_bytes_per_launch": mb * scale,
                   "source": f"profiles/{tag}_ncu_full_summary.json (ncu --set full --clock-control none, dram__bytes_read.sum + "
                             "dram__bytes_write.sum, one launch, B=64)"}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(summary["launch_list_share_of_step"], indent=1))
for d in full + cf:
    print(d["Kernel Name"][:50], d.get("gpu__time_duration.sum"), d.get("smsp__inst_executed.sum"), d.get("dram__bytes_read.sum"))
