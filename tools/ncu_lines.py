"""Aggregate an ncu source page (needs -lineinfo + --import-source on) by CUDA source line."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
extra = sys.argv[3:]   # e.g. --kernel-name regex:psn_strip --launch-skip 0 --launch-count 1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; cur_line = None; agg = collections.Counter(); inst = collections.Counter(); src = {}; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; iS = r.index("Warp Stall Sampling (All Samples)"); iE = r.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0])); src[cur_line] = r[1]
    else:
        try:
            agg[cur_line] += int(r[iS]); inst[cur_line] += int(r[iE])
        except Exception:
            pass
tot = sum(agg.values()) or 1; toti = sum(inst.values()) or 1
print("total samples", tot, "total warp-instructions", toti)
for k, v in agg.most_common(topn):
    print(f"{100*v/tot:5.1f}% samples {100*inst[k]/toti:5.1f}% inst  {k[0]}:{k[1]}  {src[k].strip()[:110]}")
