#!/usr/bin/env python
"""Summarise an ncu capture (.ncu-rep) and a launch list (csv) into profiles/<tag>.md (+ spmv_traffic.json).

    python scripts/ncu_summary.py <tag> [--rep gpurun_out/prof_<tag>.ncu-rep] [--launches gpurun_out/launches_<tag>.csv]

Runs here (no GPU needed): `ncu -i` only reads the report."""
import argparse
import collections
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * mult.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--traffic-kernel", default="spmv", help="kernel-name substring whose DRAM bytes go to spmv_traffic.json")
    a = ap.parse_args()
    rep = a.rep or os.path.join(ROOT, "gpurun_out", f"prof_{a.tag}.ncu-rep")
    launches = a.launches or os.path.join(ROOT, "gpurun_out", f"launches_{a.tag}.csv")
    out = [f"# ncu summary `{a.tag}`", ""]
    if os.path.exists(launches):
        per = collections.OrderedDict()
        with open(launches) as f:
            lines = [l for l in f if l.startswith('"')]
        for row in csv.DictReader(io.StringIO("".join(lines))):
            if row.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = row["Kernel Name"].split("(")[0]
            per.setdefault(name, []).append(float(row["Metric Value"]) * (1e-3 if row["Metric Unit"] in ("ns", "nsecond") else 1.0))
        total = sum(sum(v) for v in per.values())
        out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)",
                "", "| kernel | launches | total us | mean us | share |", "|---|---:|---:|---:|---:|"]
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            out.append(f"| `{name}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.2f} | {100 * sum(v) / total:.1f}% |")
        top = max(per.items(), key=lambda kv: sum(kv[1]))
        big = [v for v in top[1] if v > 0.5 * max(top[1])]
        out += ["", f"`{top[0]}`: {len(big)} whole-matrix launches (the warm-up + timed steps of the bench), mean {sum(big) / len(big):.1f} us each; "
                    f"the other {len(top[1]) - len(big)} launches are the row chunks of the host-buffer (e2e) path and per-chunk plan kernels. "
                    "Inside the timed region the step IS this one launch (share 100 %)."]
        out.append("")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        out += ["## Full capture (`ncu --set full --clock-control none --import-source on`)", ""]
        for r in data:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            out.append(f"### `{d['Kernel Name'].split('(')[0]}`  grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}")
            out += ["", "| metric | value | unit |", "|---|---:|---|"]
            for k in KEYS:
                if k in d:
                    out.append(f"| {k} | {d[k]} | {u[k]} |")
            stalls = []
            for h in hdr:
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                    try:
                        stalls.append((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            out.append("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + " | |")
            out.append("")
            if a.traffic_kernel in d["Kernel Name"]:
                rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
                wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
                with open(os.path.join(ROOT, "profiles", "spmv_traffic.json"), "w") as f:
                    json.dump({"kernel": d["Kernel Name"].split("(")[0], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd,
                               "dram_bytes_write": wr, "source": f"profiles/{a.tag}.md (ncu --set full, one launch)"}, f, indent=1)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{a.tag}.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
