#!/usr/bin/env python
"""(build container) Per-kernel counts of the Blackwell-specific SASS mnemonics in csrc/libms_b200.so, from
``cuobjdump -sass``: UTCIMMA/UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UBLKCP (bulk copy), LDTM (tcgen05.ld),
SYNCS (mbarrier), FADD2/FMUL2/FFMA2 (packed fp32), plus the register count and static shared memory from ``cuobjdump -res-usage``.  Writes
profiles/<round>_sass_summary.txt stamped with the hash of the sources the library was built from.

    python tools/sass_summary.py [r02]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meteor_scatter_b200 import build as _build   # noqa: E402

MNEMONICS = ("UTCIMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "SYNCS", "UTCBAR",
             "HMMA", "IMMA", "DFMA", "DMUL", "DADD", "MUFU", "FADD2", "FMUL2", "FFMA2")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    so = _build.LIB_PATH
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function ([^:]+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    kernels[cur][mn] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = [f"# SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass), source hash {_build.source_hash()}",
           "# columns: instructions, registers, static smem, then non-zero counts of " + " ".join(MNEMONICS), ""]
    for (k, c), name in zip(kernels.items(), demangle):
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        reg, sh = usage.get(k, (None, None))
        cnt = " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
        out.append(f"{name:<58s} instr={c['_total']:<6d} reg={reg} smem={sh}  {cnt}")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
