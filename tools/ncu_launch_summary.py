"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log into the
per-kernel launch table kept under profiles/ (and refresh profiles/ncu_traffic.json).

usage: python tools/ncu_launch_summary.py gpurun_out/launches.csv "<command that was profiled>" [round-tag]
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

STAGE_OF = [("preprocess_kernel", "preprocess"), ("presort_keys_kernel", "presort"), ("onesweep_kernel<unsigned long", "presort"),
            ("scan_kernel", "scan"), ("duplicate_coop_kernel", "duplicate"), ("duplicate_kernel", "duplicate"),
            ("onesweep_kernel", "sort"), ("hist_kernel", "sort"), ("tile_ranges", "ranges"),
            ("render_fwd", "render_fwd"), ("render_bwd", "render_bwd"), ("backward_gaussians", "gauss_bwd")]


def main():
    path, cmd = sys.argv[1], sys.argv[2]
    tag = sys.argv[3] if len(sys.argv) > 3 else "r1"
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = OrderedDict()
    for r in rows:
        name = re.sub(r"^void ", "", r["Kernel Name"]).replace("<unnamed>::", "")
        name = re.sub(r"\(.*$", "", name)
        d = per.setdefault(name, {"ids": set(), "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        d["ids"].add(r["ID"])
        if m == "gpu__time_duration.sum":
            d["us"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["rd" if "read" in m else "wr"] += b
    total = sum(d["us"] for d in per.values())
    out = [f"# ncu launch list, {tag} (C2: 1M Gaussians, SH3, 1920x1088, :rgbd, default strict math)",
           f"# command: {cmd}",
           "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's `stages`, not absolutes",
           f"{'kernel':44s} {'launches':>8s} {'avg_us':>9s} {'share%':>7s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s}"]
    traffic = {}
    for name, d in per.items():
        n = len(d["ids"])
        out.append(f"{name[:44]:44s} {n:8d} {d['us'] / n:9.1f} {100 * d['us'] / total:7.1f} {d['rd'] / n / 1e6:11.1f} {d['wr'] / n / 1e6:11.1f}")
        for key, stage in STAGE_OF:
            if key in name:
                t = traffic.setdefault(stage, [0.0, 0])
                t[0] += d["rd"] + d["wr"]
                t[1] = max(t[1], n)
                break
    txt = "\n".join(out) + "\n"
    dst = os.path.join("profiles", f"{tag}_ncu_launches.txt")
    open(dst, "w").write(txt)
    print(txt)
    # per-step traffic of a stage = all of its launches in one step; steps = launches of a once-per-step kernel
    steps = max(1, min(v[1] for k, v in traffic.items() if k in ("render_fwd", "render_bwd")))
    tj = {"source": f"{dst} (ncu dram__bytes_read.sum + dram__bytes_write.sum, summed over a stage's launches, per step)",
          "per_launch_bytes": {k: int(v[0] / steps) for k, v in traffic.items()}}
    json.dump(tj, open(os.path.join("profiles", "ncu_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
