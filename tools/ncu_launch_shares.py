"""Per-kernel share of the device time in an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`), with the
strip kernels split by leg (pre / post) and level (0: double rhs, coarse: float rhs).  Compare with the CUDA-event shares
in bench.py's "kernels" object: ncu times are cold-cache and serialised, the SHARES must agree.
   python tools/ncu_launch_shares.py profiles/r2_ncu_launch_list_bench.csv [bench.json]"""
import collections, csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg, cnt = collections.Counter(), collections.Counter()
ALIAS = {"psn_upre_strip_kernel": "update_pre_l0", "psn_dir_strip_kernel": "dir_apply_dot", "psn_flush_x_kernel": "flush_x",
         "psn_prologue_kernel": "prologue", "psn_true_residual_strip_kernel": "true_residual", "psn_coarse_vcycle_kernel": "coarse_vcycle", "psn_block_vcycle_kernel": "coarse_vcycle",
         "psn_update_strip_kernel": "update_norm", "psn_block_solve_kernel": "smem_solve", "psn_smem_solve_kernel": "smem_solve"}
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    name, val, unit = r[hdr.index("Kernel Name")], r[hdr.index("Metric Value")], r[hdr.index("Metric Unit")]
    try:
        v = float(val.replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    m = re.search(r"(psn_\w+)(?:<([^>]*)>)?", name)
    if not m:
        k = "other (torch / memset)"
    elif m.group(1) == "psn_strip_kernel":
        t = [x.strip() for x in m.group(2).split(",")]
        k = ("strip_post" if t[1] == "1" else "strip_pre") + ("_l0" if t[3] == "double" else "_coarse")
    elif m.group(1) == "psn_upre_strip_kernel":
        t = [x.strip() for x in (m.group(2) or "2, 0").split(",")]
        k = "prologue_pre_l0" if t[-1] == "1" else "update_pre_l0"
    elif m.group(1).startswith("psn_state") or m.group(1) in ("psn_publish_counter_kernel", "psn_finalize_kernel", "psn_zero_ints_kernel"):
        k = "state"
    else:
        k = ALIAS.get(m.group(1), m.group(1))
    agg[k] += v
    cnt[k] += 1
ev = {}
if len(sys.argv) > 2:
    ev = json.loads(open(sys.argv[2]).read()).get("kernels", {})
tot = sum(v for k, v in agg.items() if k in ev or not ev)
print(f"{'kernel':24s} {'launches':>8s} {'ncu ms':>10s} {'ncu share':>10s} {'CUDA-event share (bench.py)':>28s}")
for k, v in agg.most_common():
    print(f"{k:24s} {cnt[k]:8d} {v:10.3f} {100 * v / tot:9.1f}% {('%.1f%%' % (100 * ev[k]['share'])) if k in ev else '':>28s}")
