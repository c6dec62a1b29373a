#!/usr/bin/env python
"""Summarise the SASS of libsmb200.so per kernel family into profiles/sass_<tag>.md: the mnemonics that prove the
Blackwell-native paths (UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier, LDG...256, system-scope loads/stores of
the peer-memory protocol) plus registers, static shared memory and code size.  Runs without a GPU (cuobjdump only).

    python scripts/sass_summary.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparsemat_b200", "lib", "libsmb200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda names: dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))

usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(v) for v in m.groups())

PATTERNS = collections.OrderedDict([
    ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDG.256", r"\bLDG\.E\S*\.256"), ("LDG.128", r"\bLDG\.E\S*\.128"),
    ("LDS", r"\bLDS"), ("LD/ST .SYS", r"\b(LD|ST|LDG|STG)\.E\S*\.SYS"), ("MEMBAR.SYS", r"\bMEMBAR\S*\.SYS|\bFENCE\S*SYS"),
    ("ATOM/RED", r"\b(ATOMG|ATOM|RED|REDG|ATOMS)\b"), ("MATCH", r"\bMATCH"), ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\."),
    ("HMMA/UTC*MMA", r"\b(HMMA|UTC\w*MMA|LDTM|STTM)"),
])
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if cur and m:
        per[cur]["_n"] += 1
        for name, pat in PATTERNS.items():
            if re.search(pat, m.group(1)):
                per[cur][name] += 1

names = demangle(list(per))
fam = collections.OrderedDict()
for k, c in per.items():
    d = names.get(k, k)
    base = re.sub(r"^void ", "", d)
    base = re.sub(r"\(.*", "", base)
    family = re.sub(r"<.*", "", base).replace("smb::", "")
    fam.setdefault(family, []).append((base.replace("smb::", ""), c, usage.get(k, (0, 0, 0))))

out = [f"# SASS summary `{tag}` — sparsemat_b200/lib/libsmb200.so (sm_100a only)", "",
       "`cuobjdump -sass` / `--dump-resource-usage`, counted per kernel; one row per kernel family (ranges over its template",
       "instantiations).  UBLKCP = `cp.async.bulk` (TMA bulk copy), SYNCS = mbarrier operations, `LD/ST .SYS` = the",
       "system-scope acquire loads / release stores of the peer-memory protocol (halo.cuh).  No tensor-core instruction",
       "anywhere: SpMV is a bandwidth-bound gather (BASELINE.json north_star).", "",
       "| kernel family | inst. | SASS instr. | regs | static smem | " + " | ".join(PATTERNS) + " |",
       "|---|---:|---:|---:|---:|" + "---:|" * len(PATTERNS)]


def rng(vals):
    lo, hi = min(vals), max(vals)
    return str(lo) if lo == hi else f"{lo}–{hi}"


for family, items in sorted(fam.items(), key=lambda kv: -max(c["_n"] for _, c, _ in kv[1])):
    cols = [rng([c[p] for _, c, _ in items]) for p in PATTERNS]
    out.append(f"| `{family}` | {len(items)} | {rng([c['_n'] for _, c, _ in items])} | {rng([u[0] for _, _, u in items])} | "
               f"{rng([u[2] for _, _, u in items])} | " + " | ".join(cols) + " |")
spills = [(n, u) for items in fam.values() for n, _, u in items if u[1]]
out += ["", f"Kernels: {sum(len(v) for v in fam.values())} in {len(fam)} families.  "
        f"Kernels with a local-memory stack frame: {len(spills)}" + (": " + ", ".join(f"`{n[:60]}` ({u[1]} B)" for n, u in spills[:12]) if spills else "") + ".", ""]
ring = [(n, c, u) for n, c, u in fam.get("spmv_ring_kernel", [])]
if ring:
    out += ["## spmv_ring_kernel instantiations (the headline kernel; last template flag = distributed launch)", "",
            "| instantiation | SASS instr. | code KB | regs | UBLKCP | SYNCS | LD/ST .SYS |", "|---|---:|---:|---:|---:|---:|---:|"]
    for n, c, u in ring:
        out.append(f"| `{n}` | {c['_n']} | {c['_n'] * 16 / 1024:.1f} | {u[0]} | {c['UBLKCP']} | {c['SYNCS']} | {c['LD/ST .SYS']} |")
path = os.path.join(ROOT, "profiles", f"sass_{tag}.md")
with open(path, "w") as f:
    f.write("\n".join(out) + "\n")
print(path)
