#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/:
  python tools/summarize_ncu.py <round> <launches.csv> <full.ncu-rep>
-> profiles/launches_<round>.md (+ the csv), profiles/ncu_<round>.json / .md (one entry per captured kernel)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

def kname(s):
    """'void k_presiso<0, 1>(const float2 *, ...)' -> 'k_presiso<0, 1>'"""
    s = s.split("(")[0].replace("<unnamed>::", "").strip()
    return s[5:] if s.startswith("void ") else s


rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = kname(r[4])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1].replace(",", ""))
SETUP = ("k_tx_",)                                              # input synthesis: launched before the timed region
tot = sum(v[1] for k, v in agg.items() if k.startswith("k_") and not k.startswith(SETUP))
with open(os.path.join(out, "launches_%s.md" % rnd), "w") as f:
    f.write("# ncu launch list, %s (`--metrics gpu__time_duration.sum --clock-control none`, bench.py --frames 113664 --steps 2 --warmup 1 --no-cpu: two chunks per step)\n\n" % rnd)
    f.write("Per-launch times are cold-cache and serialised by ncu: compare SHARES with bench.py's `stages`, not absolutes.\n\n")
    f.write("| kernel | launches | total ms | avg ms | share of the step's kernels |\n|---|---|---|---|---|\n")
    for k, v in agg.items():
        if k.startswith("k_"):
            share = "(input synthesis, outside the timed region)" if k.startswith(SETUP) else "%.3f" % (v[1] / tot)
            f.write("| %s | %d | %.3f | %.3f | %s |\n" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e6, share))
    f.write("\n(torch kernels that build the synthetic batch are in the csv but not listed here.)\n")
subprocess.call(["cp", launches, os.path.join(out, "launches_%s.csv" % rnd)])

raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rr[0], rr[1], rr[2:]
want = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_sm",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wavefronts_pct",
    "smsp__inst_issued.sum": "warp_instructions",
    "smsp__average_warp_latency_per_inst_issued.ratio": "warp_latency_per_inst",
}
kernels = []
for d in data:
    k = {"kernel": kname(d[hdr.index("Kernel Name")])}
    for m, short in want.items():
        if m in hdr:
            k[short] = "%s %s" % (d[hdr.index(m)], units[hdr.index(m)])
    st = {}
    for i, h in enumerate(hdr):
        if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(float(d[i]), 3)
            except ValueError:
                pass
    k["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
    kernels.append(k)
json.dump(kernels, open(os.path.join(out, "ncu_%s.json" % rnd), "w"), indent=1)
with open(os.path.join(out, "ncu_%s.md" % rnd), "w") as f:
    f.write("# ncu --set full --clock-control none, %s (one launch per kernel, bench.py --frames 56832: one chunk of config-5 items)\n\n" % rnd)
    for k in kernels:
        f.write("## %s\n\n" % k["kernel"])
        for kk, vv in k.items():
            if kk != "kernel":
                f.write("- %s: %s\n" % (kk, vv))
        f.write("\n")
print("wrote profiles/ for", rnd, [k["kernel"] for k in kernels])
