#!/usr/bin/env python
"""Summarise ONE `ncu --set full --import-source on` capture as markdown on stdout (run on the GPU box, right after the
capture, so that only text travels back).  Usage: summarize_one.py <file.ncu-rep> <name>"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import summarize_profiles as sp  # noqa: E402


def main():
    rep, name = sys.argv[1], sys.argv[2]
    m = sp.raw_page(rep)
    out = [f"# r2: `ncu --set full --clock-control none --import-source on` capture, kernel `{name}`\n", "| metric | value | unit |", "|---|---:|---|"]
    for k in sp.KEY_METRICS + ["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]:
        if k in m:
            out.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    out.append("\nWarp stall reasons (warps per issue-active cycle):\n")
    out += ["| reason | value |", "|---|---:|"]
    st = [(k.split("issue_stalled_")[1].split("_per_issue")[0], float(v[0])) for k, v in m.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for k, v in sorted(st, key=lambda x: -x[1]):
        if v > 0.005:
            out.append(f"| {k} | {v:.3f} |")
    src = sp.source_page(rep, top=40)
    if src:
        out.append("\nHottest source lines by stall samples:\n")
        out.append(src)
    print("\n".join(out))


if __name__ == "__main__":
    main()
