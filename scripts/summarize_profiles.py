#!/usr/bin/env python
"""Turn the raw captures scripts/profile.sh left in gpurun_out/ into the tracked summaries under profiles/.
Usage: python scripts/summarize_profiles.py <tag>        (needs `ncu` on PATH to read the .ncu-rep files)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEY_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
]


def launches(tag):
    src = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return None
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        f.writelines(lines)
    out = [f"# {tag}: launch list of the bench command (ncu --metrics gpu__time_duration.sum --clock-control none)\n",
           f"{len(rows) - 1} launches, {tot:.1f} ms of kernel time (cold-cache, serialised: compare SHARES, not absolutes)\n",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {k} | {a[0]} | {a[1]:.3f} | {a[1] / tot:.3f} |")
    return "\n".join(out) + "\n"


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def source_page(rep, top=30):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    agg = {}
    hdr = None
    n_sass = 0
    for r in rows:
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            il, isrc = 0, 1
            iw = hdr.index("Warp Stall Sampling (All Samples)")
            ii = hdr.index("Instructions Executed")
            ilsb = hdr.index("stall_long_sb") if "stall_long_sb" in hdr else None
            ini = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        try:
            ie, ws = int(r[ii] or 0), int(r[iw] or 0)
        except ValueError:
            continue
        if r[il] == "":       # SASS view rows (no line number): count them, the per-line view carries the samples
            n_sass += 1
            continue
        a = agg.setdefault((r[il], r[isrc].strip()), [0, 0, 0, 0, 0])
        a[0] += ie
        a[1] += ws
        a[2] += 1
        a[3] += int(r[ilsb] or 0) if ilsb is not None else 0
        a[4] += int(r[ini] or 0) if ini is not None else 0
    if not agg:
        return None
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    out = [f"{n_sass} SASS instructions in the kernel; shares are of all executed warp instructions / all stall samples.\n",
           "| line | SASS | instr share | stall share | long_sb | no_inst | source |", "|---:|---:|---:|---:|---:|---:|---|"]
    for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"| {ln} | {a[2]} | {a[0] / tot_i:.3f} | {a[1] / tot_s:.3f} | {a[3]} | {a[4]} | `{src[:100]}` |")
    return "\n".join(out) + "\n"


def kernel_summary(tag, name):
    rep = os.path.join(OUT, f"prof_{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return None, None
    m = raw_page(rep)
    out = [f"# {tag}: `ncu --set full --clock-control none --import-source on` capture, kernel `{name}`\n",
           "| metric | value | unit |", "|---|---:|---|"]
    for k in KEY_METRICS:
        if k in m:
            out.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    out.append("\nWarp stall reasons (warps per issue-active cycle):\n")
    out += ["| reason | value |", "|---|---:|"]
    st = [(k.split("issue_stalled_")[1].split("_per_issue")[0], float(v[0])) for k, v in m.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for k, v in sorted(st, key=lambda x: -x[1]):
        if v > 0.005:
            out.append(f"| {k} | {v:.3f} |")
    src = source_page(rep)
    if src:
        out.append("\nHottest source lines by stall samples:\n")
        out.append(src)
    def num(k):
        v, u = m[k]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    facts = {"kernel": name, "duration_ms": float(m["gpu__time_duration.sum"][0]) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[m["gpu__time_duration.sum"][1]],
             "dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
             "warp_instructions": float(m["smsp__inst_executed.sum"][0])}
    return "\n".join(out) + "\n", facts


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    s = launches(tag)
    if s:
        open(os.path.join(PROF, f"{tag}_launches_summary.md"), "w").write(s)
    facts = {}
    for name in ("seq", "decode"):
        md, f = kernel_summary(tag, name)
        if md:
            open(os.path.join(PROF, f"{tag}_ncu_{name}.md"), "w").write(md)
            facts[name] = f
    for extra in (f"bench_{tag}_profcfg.json", f"host_{tag}.txt"):
        p = os.path.join(OUT, extra)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, f"{tag}_{extra.replace('_' + tag, '')}"))
    p = os.path.join(PROF, f"{tag}_bench_profcfg.json")
    if os.path.exists(p) and facts:
        line = json.loads(open(p).read().strip().splitlines()[-1])
        rows = line["config"]["op_rows_per_gpu"]
        for f in facts.values():
            f["op_rows_per_launch"] = rows
            f["dram_bytes_per_op_row"] = f["dram_bytes"] / rows
            f["warp_instructions_per_op_row"] = f["warp_instructions"] / rows
    json.dump(facts, open(os.path.join(PROF, f"{tag}_ncu_facts.json"), "w"), indent=1)
    print(json.dumps(facts, indent=1))


if __name__ == "__main__":
    main()
