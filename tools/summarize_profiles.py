#!/usr/bin/env python3
"""Turn ncu outputs into the tracked summaries under profiles/.

  summarize_profiles.py TAG LAUNCHES.csv FULL_RAW.csv [N_REGULAR N_PML N_NODES [N_PML_NODES]]
  summarize_profiles.py TAG LAUNCHES.csv FULL_RAW.csv kernel=work_items ... [case=bench.tpv104_100m]

LAUNCHES.csv: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list.
FULL_RAW.csv: `ncu -i prof.ncu-rep --page raw --csv` of a `--set full` capture.
Writes profiles/ncu_summary_TAG.json and refreshes profiles/ncu_summary.json
(read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    return name.replace("void ", "").split("(")[0].split("<")[0].replace("eqd::", "")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        val = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        val_us = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += val_us
    tot = sum(v[1] for v in agg.values())
    return {k: {"launches": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 1), "share": round(v[1] / tot, 4)}
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}


KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def to_bytes(val, unit):
    m = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * m.get(unit, 1.0)


def to_ms(val, unit):
    m = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    return float(val.replace(",", "")) * m.get(unit, 1.0)


def full(path, counts):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])   # a later launch of the same kernel replaces an earlier one: the steady-state step
        d = {"template": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for k in KEEP:
            if k in hdr:
                d[k] = (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip()
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        ms = to_ms(r[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
        d["dram_bytes_per_launch"] = rd + wr
        d["dram_GBps_under_ncu"] = round((rd + wr) / (ms * 1e-3) / 1e9, 1)
        n = counts.get(name)
        if n:
            d["work_items"] = n
            d["dram_bytes_per_element"] = round((rd + wr) / n, 1)
        out[name] = d
    return out


def main():
    tag, lpath, fpath = sys.argv[1:4]
    counts = {}
    case = None
    rest = sys.argv[4:]
    if rest and all("=" in a for a in rest):       # kernel=work_items ... [case=NAME]
        for a in rest:
            k, v = a.split("=", 1)
            if k == "case":
                case = v
            else:
                counts[k] = int(v)
    elif len(rest) >= 3:
        counts = {"k_tile_reg": int(rest[0]), "k_tile_pml": int(rest[1]), "k_node_update3": int(rest[2])}
        if len(rest) >= 4:
            counts["k_node_update12"] = int(rest[3])
    s = {"tag": tag, "case": case, "launch_list": launches(lpath), **full(fpath, counts)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for name in ("ncu_summary_%s.json" % tag, "ncu_summary.json"):
        json.dump(s, open(os.path.join(ROOT, "profiles", name), "w"), indent=1)
    print(json.dumps(s, indent=1))


if __name__ == "__main__":
    main()
