#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libbasq_b200.so (evidence for profiles/: which kernels issue
tcgen05 MMAs (UTC*MMA), TMEM loads/stores (LDTM/STTM), bulk copies (UBLKCP / UTMALDG), fp64 tensor
MMAs (DMMA), MUFU.EX2, DFMA ...).  Runs on a CPU-only box: cuobjdump only reads the binary.

    python scripts/sass_histogram.py [--lib PATH] [--out profiles/rNN_sass_opcodes.md]
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS",
        "DMMA", "HMMA", "MUFU.EX2", "MUFU", "DFMA", "FFMA", "LDGSTS", "REDUX", "LDG", "STG", "ATOM", "BAR"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "basq_b200", "libbasq_b200.so"))
    ap.add_argument("--out", default="")
    ap.add_argument("--min", type=int, default=1, help="only kernels with at least this many tensor/TMEM/bulk opcodes, or MUFU/DFMA heavy")
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
    kern, hist = None, collections.OrderedDict()
    op_re = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            kern = m.group(1)
            hist[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        m = op_re.match(ln)
        if not m:
            continue
        op = m.group(1)
        hist[kern]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA") and op.startswith(k)):
                hist[kern][k] += 1
    try:
        names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    except Exception:
        names = list(hist)
    cols = [k for k in KEYS if any(h[k] for h in hist.values())]
    lines = ["| kernel | instr | " + " | ".join(cols) + " |", "|---|---:|" + "---:|" * len(cols)]
    tot = collections.Counter()
    for (k, h), nm in zip(hist.items(), names):
        for c in cols + ["_total"]:
            tot[c] += h[c]
        nm = nm.replace("(anonymous namespace)::", "").replace("basq::", "")
        nm = re.sub(r"\(.*\)$", "", re.sub(r"^void ", "", nm))
        lines.append(f"| `{nm[:70]}` | {h['_total']} | " + " | ".join(str(h[c]) if h[c] else "" for c in cols) + " |")
    lines.append(f"| **all {len(hist)} kernels** | {tot['_total']} | " + " | ".join(str(tot[c]) for c in cols) + " |")
    text = ("# SASS opcode histogram of `basq_b200/libbasq_b200.so` (cuobjdump -sass, sm_100a)\n\n"
            "`UTC*MMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier ops, "
            "`DMMA` = fp64 mma.sync, `LDGSTS` = cp.async.  Static instruction counts per kernel (not dynamic).\n\n"
            + "\n".join(lines) + "\n")
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    sys.stdout.write(text if not a.out else f"wrote {a.out}: {len(hist)} kernels\n")


if __name__ == "__main__":
    main()
