#!/usr/bin/env python
"""Turn an `ncu --set full` report of one step into the `ncu_traffic_<cfg>.json` that bench.py quotes as `roofline.traffic`.

    python profiles/make_traffic_json.py gpurun_out/fused128_C2.ncu-rep C2 256 auto k_backward > profiles/r02/ncu_traffic_C2.json

The JSON is stamped with the content hash of ptyrad_b200/csrc (bench.py: csrc_hash): bench.py refuses it for any other build of the
kernels, batch size or path.  Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed).
"""
import csv
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def short_name(name):
    """'void ptyb::fused128::k_forward<(bool)0, (bool)0>(ptyb::fused128::Args)' -> 'k_forward'; 'ptyb::k_adam(ptyb::AdamArgs)' -> 'k_adam'"""
    prev = None
    while prev != name:                                  # template arguments, innermost first
        prev, name = name, re.sub(r"<[^<>]*>", "", name)
    name = name.split("(")[0].strip()
    return name.split("::")[-1].split(" ")[-1]


def main():
    rep, cfg, batch, path, dominant = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    how = sys.argv[6] if len(sys.argv) > 6 else "ncu --set full --clock-control none"
    import bench
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def gb(row, name):
        v, u = float(row[col[name]].replace(",", "")), units[col[name]]
        return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[u]

    def ms(row):
        v, u = float(row[col["gpu__time_duration.sum"]].replace(",", "")), units[col["gpu__time_duration.sum"]]
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]

    kernels = {}
    sections = {}                                       # general path: all launches of a section (one step) summed
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        short = short_name(name)
        kernels.setdefault(short, dict(dram_read_gb=round(gb(r, "dram__bytes_read.sum"), 6), dram_write_gb=round(gb(r, "dram__bytes_write.sum"), 6),
                                       duration_ms=round(ms(r), 6)))
        sec = "adjoint_section" if short.startswith("k_bwd") else ("forward_section" if short.startswith(("k_fwd", "k_init_shift")) else None)
        if sec and dominant.endswith("_section"):
            a = sections.setdefault(sec, dict(dram_read_gb=0.0, dram_write_gb=0.0, duration_ms=0.0, launches=0))
            a["dram_read_gb"] += gb(r, "dram__bytes_read.sum"); a["dram_write_gb"] += gb(r, "dram__bytes_write.sum")
            a["duration_ms"] += ms(r); a["launches"] += 1
    for k, a in sections.items():
        kernels[k] = {kk: (round(v, 6) if isinstance(v, float) else v) for kk, v in a.items()}
    json.dump(dict(source=f"{how}, tools/prof_step.py {cfg} {path}, one launch each ({os.path.basename(rep)})",
                   csrc_sha256=bench.csrc_hash(), batch=batch, path=path, dominant=dominant, kernels=kernels), sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
