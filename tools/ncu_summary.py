#!/usr/bin/env python
"""Joins an `ncu --page raw --csv` export of tests/manual/one_step.py with the library's `dg-trace` lines (DG_TRACE_LAUNCHES=1)
and writes, per tag bench.py reports: launches, time, tensor-pipe activity, DRAM bytes, L2->SM traffic, issue activity.

    python tools/ncu_summary.py raw.csv trace.log out_prefix      ->  out_prefix.md, out_prefix.json (traffic table)

Times under the profiler are cold-cache and serialised: compare shares, not absolutes."""
import csv
import json
import sys
from collections import OrderedDict

raw, trace, prefix = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


units = rows[1]
data = [r for r in rows[2:] if len(r) > col["Kernel Name"]]
# the library's kernels in launch order (torch's own kernels -- copies, fills -- are not counted by DG_LAUNCHED)
lib = [r for r in data if not r[col["Kernel Name"]].startswith(("void at::", "at::", "void at_cuda", "void c10"))]
tags = []
inside = False
for line in open(trace):
    if line.startswith("dg-trace-begin"):
        inside, tags = True, []
    elif line.startswith("dg-trace-end"):
        inside = False
    elif inside and line.startswith("dg-trace "):
        _, name, a, b = line.split()
        tags.append((name, int(a), int(b)))
if not tags:
    raise SystemExit("no dg-trace lines between dg-trace-begin / dg-trace-end")
base = min(a for _, a, _ in tags)
# innermost scope wins: sort by span length
label = {}
for name, a, b in sorted(tags, key=lambda t: -(t[2] - t[1])):
    for i in range(a - base, b - base):
        label[i] = name
M = {"time_ns": "gpu__time_duration.sum", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "dram_r": "dram__bytes_read.sum", "dram_w": "dram__bytes_write.sum", "l2_sm": "lts__t_sectors_srcunit_tex.sum",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm_thr": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "dram_thr": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"}
scale = {}
for k, m in M.items():
    u = units[col[m]] if m in col else ""
    scale[k] = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9,
                "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(u, 1.0)
agg = OrderedDict()
for i, r in enumerate(lib):
    tag = label.get(i, "(untagged)")
    a = agg.setdefault(tag, {"n": 0, "kernels": set(), "grids": set(), "t": 0.0, "tensor_t": 0.0, "issue_t": 0.0, "r": 0.0, "w": 0.0, "l2": 0.0})
    t = f(r, M["time_ns"]) * scale["time_ns"]
    a["n"] += 1
    a["kernels"].add(r[col["Kernel Name"]].split("(")[0][:48])
    a["grids"].add(r[col["Grid Size"]])
    a["t"] += t
    a["tensor_t"] += f(r, M["tensor"]) * t
    a["issue_t"] += f(r, M["issue"]) * t
    a["r"] += f(r, M["dram_r"]) * scale["dram_r"]
    a["w"] += f(r, M["dram_w"]) * scale["dram_w"]
    a["l2"] += f(r, M["l2_sm"]) * 32.0
tot = sum(a["t"] for a in agg.values())
out = ["| tag | kernel(s) | launches | grid | time / step (us) | share | tensor pipe active | issue active | DRAM read (MB) | DRAM write (MB) | L2->SM (MB) |",
       "|---|---|---:|---|---:|---:|---:|---:|---:|---:|---:|"]
traffic = {}
for tag, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    out.append(f"| `{tag}` | {', '.join(sorted(a['kernels']))} | {a['n']} | {' / '.join(sorted(a['grids']))} | {a['t'] / 1e3:.1f} | "
               f"{100 * a['t'] / tot:.1f} % | {a['tensor_t'] / max(a['t'], 1):.1f} % | {a['issue_t'] / max(a['t'], 1):.1f} % | "
               f"{a['r'] / 1e6:.1f} | {a['w'] / 1e6:.1f} | {a['l2'] / 1e6:.1f} |")
    traffic[tag] = (a["r"] + a["w"]) / a["n"]          # per launch, like bench.py's `achieved`
traffic["_step"] = sum(a["r"] + a["w"] for a in agg.values())
traffic["_launches_per_step"] = {tag: a["n"] for tag, a in agg.items()}
out.append(f"| **total** | | {sum(a['n'] for a in agg.values())} | | {tot / 1e3:.1f} | | | | "
           f"{sum(a['r'] for a in agg.values()) / 1e6:.1f} | {sum(a['w'] for a in agg.values()) / 1e6:.1f} | |")
# a reduced copy of the raw page (the columns a reviewer needs to re-derive the table; the full page has ~2400 columns)
KEEP = ("ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_fp64", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp", "smsp__warp_issue_stalled",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg")
keep = [i for i, hname in enumerate(hdr) if hname.startswith(KEEP) and not hname.endswith((".min", ".max", ".peak_sustained"))]
with open(prefix + "_raw_selected.csv", "w", newline="") as fo:
    wr = csv.writer(fo)
    wr.writerow(["tag"] + [hdr[i] for i in keep])
    wr.writerow([""] + [units[i] for i in keep])
    for i, r in enumerate(lib):
        wr.writerow([label.get(i, "(untagged)")] + [r[j] if j < len(r) else "" for j in keep])
open(prefix + ".md", "w").write("\n".join(out) + "\n")
json.dump(traffic, open(prefix + ".json", "w"), indent=1)
print("\n".join(out))
