"""Turn ncu captures in gpurun_out/ into the small, committed evidence files under profiles/.
usage: python tools/summarize_profiles.py <launches.csv> <full.ncu-rep>[,<more.ncu-rep>...] <tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1:4]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel share of the step ("-" = none, summarise the full capture only) -------------------------
rows = [] if launches == "-" else [r for r in csv.reader(open(launches)) if len(r) > 5]
if not rows:
    rows = [["Kernel Name", "Metric Value", "Metric Unit", "", "", ""], ["", "x", "ns", "", "", ""]]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("mst::", "")
    if "at::native" in r[kn] or "elementwise" in r[kn] or "cub::" in r[kn]:
        name = "torch (setup / allocator fills)"
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
unit = rows[hdr_i + 1][hdr.index("Metric Unit")]
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6}.get(unit, 1e-6)
total = sum(a[1] for a in agg.values())
if launches != "-":
    with open(os.path.join(out_dir, f"{tag}_launch_shares.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_pct"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{t * scale:.3f}", f"{100 * t / total:.2f}"])
    subprocess.run(["cp", launches, os.path.join(out_dir, f"{tag}_launches.csv")], check=True)

# ---- full capture: key metrics per kernel --------------------------------------------------------------------------
rr = []
for one in rep.split(","):
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if not rr:
        rr = part
    else:  # align columns of later reports to the first header
        hh, uu = part[0], part[1]
        fac = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0,
               "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in part[2:]:
            row = []
            for ci, c in enumerate(rr[0]):
                if c not in hh:
                    row.append("")
                    continue
                v, u0, u1 = r[hh.index(c)], rr[1][ci], uu[hh.index(c)]
                if u0 != u1 and u0 in fac and u1 in fac:
                    try:
                        v = f"{float(v.replace(',', '')) * fac[u1] / fac[u0]:.6f}"
                    except ValueError:
                        pass
                row.append(v)
            rr.append(row)
h, units = rr[0], rr[1]
keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stall = [x for x in h if x.startswith("smsp__pcsamp_warps_issue_stalled_") and not x.endswith("_not_issued")]
idx = [h.index(k) for k in keep if k in h]
with open(os.path.join(out_dir, f"{tag}_ncu_full_summary.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow([h[i] + (f" [{units[i]}]" if units[i] else "") for i in idx] + ["top stall reasons"])
    traffic = {}
    for r in rr[2:]:
        st = {s.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[h.index(s)]) for s in stall if r[h.index(s)] not in ("", "n/a")}
        tot = sum(st.values()) or 1.0
        top = "; ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6])
        w.writerow([r[i] for i in idx] + [top])
        if "gl_kernel<0, 0>" in r[h.index("Kernel Name")] or "gl_kernel<(bool)0, (bool)0>" in r[h.index("Kernel Name")]:
            def val(m):
                i = h.index(m)
                v = float(r[i].replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
            traffic = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum")}
if traffic:
    n_clips = int(os.environ.get("PROF_CLIPS", "1024"))
    tj = {"source": f"ncu --set full, gl_kernel<false,false>, {n_clips} clips x 173 frames (profiles/{tag}_ncu_full_summary.csv)",
          "gl_iteration_dram_bytes_per_clip": (traffic["dram_bytes_read"] + traffic["dram_bytes_write"]) / n_clips, **traffic}
    with open(os.path.join(out_dir, f"{tag}_traffic_smallbatch.json"), "w") as f:  # traffic.json itself comes from the bench-size capture
        json.dump(tj, f, indent=1)
print("wrote", sorted(os.listdir(out_dir)))
