#!/usr/bin/env python
"""Reads gpurun_out/*.ncu-rep / launches CSV (here, no GPU needed) and writes the tracked summaries under profiles/:
   profiles/<name>.summary.txt  key metrics + top stall reasons + opcode mix of one `ncu --set full` capture
   profiles/traffic.json        dram bytes per launch per kernel family (read by bench.py -> roofline.traffic)
   profiles/<name>.launches.txt per-kernel share of the step from a `--metrics gpu__time_duration.sum` launch list"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def summarize(rep, name, family, traffic):
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    lines = [f"# {name}: {d.get('Kernel Name', '?')}  (ncu --set full --clock-control none, one launch; cold-cache replay)"]
    for k in KEYS:
        if k in d:
            lines.append(f"{k:75s} {d[k]} {u.get(k, '')}")
    st = {k: float(v) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k and re.fullmatch(r"[0-9.]+", v or "")}
    tot = sum(st.values()) or 1
    lines.append("top stall reasons (% of samples): " + ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.1f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[1]
        ia, ie = h.index("Source"), h.index("Instructions Executed")
        op = collections.Counter()
        for r in src[2:]:
            if len(r) <= max(ia, ie):  # a second kernel's header / separator rows in a multi-kernel report
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
            op[(m.group(2).split(".")[0] if m else "?")] += int(r[ie]) if r[ie].isdigit() else 0
        tot_i = sum(op.values()) or 1
        lines.append("opcode mix (% of executed warp instructions): " + ", ".join(f"{o} {100 * c / tot_i:.1f}" for o, c in op.most_common(14)))
        sass = " ".join(op.keys())
        lines.append("TMA / mbarrier in SASS: " + ", ".join(x for x in ("UBLKCP", "SYNCS", "UTMALDG") if x in sass))
    rd = float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
    lines.append(f"dram traffic per launch: {rd + wr:.0f} bytes (read {rd:.0f} + write {wr:.0f})")
    traffic[family] = int(rd + wr)
    open(os.path.join(ROOT, "profiles", name + ".summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:6]))


def launches(csv_path, name):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 14 and r[0].isdigit()]
    by = collections.defaultdict(list)
    for r in rows:
        by[re.sub(r"\(.*", "", r[4])].append(float(r[14]))
    tot = sum(sum(v) for v in by.values())
    lines = [f"# {name}: ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare SHARES)"]
    for k, v in sorted(by.items(), key=lambda x: -sum(x[1])):
        lines.append(f"{k:60s} launches {len(v):4d}  total {sum(v) / 1e3:10.1f} us  avg {sum(v) / len(v) / 1e3:9.1f} us  share {100 * sum(v) / tot:5.1f}%")
    open(os.path.join(ROOT, "profiles", name + ".launches.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    for fam in ("scan", "build", "probe", "probe_index", "derive", "derivepart"):
        rep = os.path.join(ROOT, "gpurun_out", f"prof_{fam}_{tag}.ncu-rep")
        if os.path.exists(rep):
            summarize(rep, f"{fam}_{tag}", fam, traffic)
    json.dump(traffic, open(tp, "w"), indent=1)
    lc = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if os.path.exists(lc):
        launches(lc, f"step_{tag}")
