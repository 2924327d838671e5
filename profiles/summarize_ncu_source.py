#!/usr/bin/env python
"""Condense the per-instruction (SASS) page of an ncu report into the two tables the profile READMEs quote:
stall samples by opcode and by stall reason, per kernel.

    python profiles/summarize_ncu_source.py gpurun_out/prof.ncu-rep [kernel-regex ...] > profiles/rNN/ncu_stalls_<name>.txt

Needs the report to have been captured with `--set full --import-source on` from a `-lineinfo` build; reads it with
`ncu -i <rep> --page source --csv` (no GPU needed).
"""
import collections
import csv
import subprocess
import sys


def source_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None, None, []
    name = rows[0][1] if len(rows[0]) > 1 else kernel
    return name, rows[1], rows[2:]


def summarize(rep, kernel):
    name, hdr, rows = source_rows(rep, kernel)
    if not rows:
        print(f"== {kernel}: no such kernel in {rep}")
        return
    i_s, i_src, i_inst = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    i_a = hdr.index("Address")
    by_op, by_stall, inst, tot, n_instr = collections.Counter(), collections.Counter(), collections.Counter(), 0, 0
    seen = set()
    for r in rows:
        if len(r) < len(hdr) or not r[i_s].isdigit() or r[i_a] in seen:      # the page lists every instruction twice
            continue
        seen.add(r[i_a])
        t = r[i_src].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        n = int(r[i_s])
        by_op[op] += n
        inst[op] += int(r[i_inst] or 0)
        tot += n
        n_instr += 1
        for i in stall:
            by_stall[hdr[i]] += int(r[i] or 0)
    print(f"== {name}")
    print(f"   SASS instructions {n_instr} ({n_instr * 16 // 1024} KB), warp-stall samples {tot}")
    print("   samples by opcode:      " + ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in by_op.most_common(12)))
    print("   samples by stall reason: " + ", ".join(f"{o[6:]} {100 * c / tot:.1f}%" for o, c in by_stall.most_common(10)))
    print("   warp instructions executed: " + ", ".join(f"{o} {c / 1e6:.1f}M" for o, c in inst.most_common(10)))


if __name__ == "__main__":
    rep = sys.argv[1]
    for k in (sys.argv[2:] or ["."]):
        summarize(rep, k)
