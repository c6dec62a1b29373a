#!/usr/bin/env python
"""Summarise raw-page CSV exports of ncu captures (made on the GPU box with `ncu -i x.ncu-rep --page raw --csv`, because the
reports themselves are too large to bring back) and a launch list into profiles/<tag>.md (+ profiles/spmv_traffic.json).

    python scripts/ncu_raw_summary.py <tag> --raw name=path.csv [--raw ...] [--launches launches.csv] [--traffic name]"""
import argparse
import collections
import csv
import io
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
]
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--raw", action="append", default=[])
    ap.add_argument("--launches")
    ap.add_argument("--traffic")
    ap.add_argument("--note", action="append", default=[])
    a = ap.parse_args()
    out = [f"# ncu summary `{a.tag}`", ""] + [n for n in a.note] + ([""] if a.note else [])
    if a.launches and os.path.exists(a.launches):
        per = collections.OrderedDict()
        with open(a.launches) as f:
            lines = [l for l in f if l.startswith('"')]
        for row in csv.DictReader(io.StringIO("".join(lines))):
            if row.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = row["Kernel Name"].split("(")[0]
            v = float(row["Metric Value"].replace(",", ""))
            per.setdefault(name, []).append(v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(row["Metric Unit"], 1.0))
        total = sum(sum(v) for v in per.values())
        out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
                "| kernel | launches | total us | mean us | share |", "|---|---:|---:|---:|---:|"]
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            out.append(f"| `{name}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.2f} | {100 * sum(v) / total:.1f}% |")
        out.append("")
    traffic = None
    for spec in a.raw:
        name, path = spec.split("=", 1)
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        out += [f"## `{name}` — {len(data)} launch(es) of `{data[0][idx['Kernel Name']][:110]}`", "", "| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " | unit |",
                "|---|" + "---:|" * len(data) + "---|"]
        for k in KEYS:
            if k in idx:
                out.append(f"| {k} | " + " | ".join(r[idx[k]] for r in data) + f" | {units[idx[k]]} |")
        r = data[0]
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i])) for h, i in idx.items()
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[i] not in ("", "n/a")]
        tot = sum(v for _, v in st) or 1.0
        out.append("| stall samples (launch 0) | " + ", ".join(f"{h} {100 * v / tot:.1f}%" for h, v in sorted(st, key=lambda t: -t[1])[:7]) + " |" * len(data) + " |")
        out.append("")
        if a.traffic == name:
            rd = float(r[idx["dram__bytes_read.sum"]]) * MULT.get(units[idx["dram__bytes_read.sum"]], 1)
            wr = float(r[idx["dram__bytes_write.sum"]]) * MULT.get(units[idx["dram__bytes_write.sum"]], 1)
            traffic = {"kernel": r[idx["Kernel Name"]][:80], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "source": f"profiles/{a.tag}.md (ncu --set full, launch 0 of `{name}`)"}
    with open(os.path.join(ROOT, "profiles", f"{a.tag}.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    if traffic:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    print("wrote profiles/%s.md" % a.tag)


if __name__ == "__main__":
    main()
