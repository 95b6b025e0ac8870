"""Developer tool: per-CUDA-source-line executed warp instructions from an ncu report with -lineinfo / --import-source.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [kernel-substring] [top-n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = None; hdr = None; cur = None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, ""]))
files = {}
fpath = None
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iN = hdr.index("# Samples"); continue
    if hdr is None: continue
    if r[0] != "": cur = (fpath.split("/")[-1], int(r[0])); agg[fn][cur][2] = r[1].strip(); continue
    try: n = int(r[iI]); s = int(r[iN])
    except ValueError: continue
    agg[fn][cur][0] += n; agg[fn][cur][1] += s
for fn, lines in agg.items():
    if want not in fn: continue
    tot = sum(v[0] for v in lines.values()); smp = sum(v[1] for v in lines.values())
    print(f"===== {fn[:80]}  inst={tot} samples={smp}")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]:9d} {100*v[0]/max(tot,1):5.1f}%  smp {100*v[1]/max(smp,1):5.1f}%  {f}:{ln:<4d} {v[2][:100]}")
