#!/usr/bin/env python
"""tools/ncu_summary.py output -> one line per profiled launch: duration, DRAM bytes (read + write), achieved DRAM GB/s and its
fraction of the measured HBM peak (MEASURED_PEAKS.json), registers, achieved occupancy, issue-slot utilisation, long-scoreboard
stall.  `python tools/ncu_table.py summary.txt [--json out.json] [--labels "a|b|c"]`"""
import argparse
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def parse(path):
    rows, cur = [], None
    for line in open(path):
        if line.startswith("=="):
            cur = {"kernel": line[3:].strip()}
            rows.append(cur)
        elif cur is not None:
            m = re.match(r"\s+(\S+)\s+(\S+)\s*(\S*)", line)
            if m:
                try:
                    cur[m.group(1)] = (float(m.group(2).replace(",", "")), m.group(3))
                except ValueError:
                    pass
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("summary")
    ap.add_argument("--json", default=None)
    ap.add_argument("--labels", default=None, help="'|'-separated labels, one per launch, replacing the kernel names in the JSON")
    args = ap.parse_args()
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    labels = args.labels.split("|") if args.labels else None
    out = []
    for i, r in enumerate(parse(args.summary)):
        def val(key, table=None, default=0.0):
            if key not in r:
                return default
            v, u = r[key]
            return v * table.get(u, 1.0) if table else v
        us = val("gpu__time_duration.sum", TIME)
        rd, wr = val("dram__bytes_read.sum", UNIT), val("dram__bytes_write.sum", UNIT)
        gbs = (rd + wr) / us / 1e3 if us else 0.0
        name = re.sub(r"^void ", "", r["kernel"])
        name = re.sub(r"unnamed>::", "", name)
        name = re.sub(r"\(.*$", "", name)
        row = {"kernel": name, "label": labels[i] if labels and i < len(labels) else None, "duration_us": us,
               "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_GBs": gbs, "frac_of_hbm_peak": gbs / peak,
               "registers": int(val("launch__registers_per_thread")),
               "achieved_occupancy_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
               "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "stall_long_scoreboard": val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")}
        out.append(row)
        print("%-58s %8.1f us %9.1f MB %7.0f GB/s %5.1f%%  regs %3d  occ %4.1f%%  issue %4.1f%%" % (
            (row["label"] or name)[:58], us, (rd + wr) / 1e6, gbs, 100 * gbs / peak, row["registers"],
            row["achieved_occupancy_pct"], row["issue_active_pct"]))
    if args.json:
        json.dump({"hbm_peak_gbs": peak, "source": os.path.basename(args.summary),
                   "what": "per-launch figures from one `ncu --set full --clock-control none` capture (cold-ish caches, serialised "
                           "launches: durations are upper bounds of the in-step times, DRAM bytes are a property of the access pattern)",
                   "rows": out}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
