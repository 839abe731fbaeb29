#!/usr/bin/env python
"""Static evidence for the shipped kernels (no GPU needed): per kernel registers / spills / shared memory from
`cuobjdump -res-usage` and the count of the SASS mnemonics that prove the tensor-core / async-copy paths
(/opt/skills/guides/B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async -> LDGSTS,
mbarrier -> SYNCS, packed fp32 -> FFMA2).

    python tools/sass_summary.py > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "echoglad_b200", "libechoglad_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS",
        "FENCE.VIEW.ASYNC", "FFMA2", "FADD2", "FMUL2", "LDG", "STG", "LDS", "STS", "SHFL", "LDL", "STL", "DADD",
        "USETMAXREG", "NANOSLEEP"]


def short(name):
    base = name
    end = name.find("_kernel")
    if end >= 0:
        end += len("_kernel")
        for n in range(4, 40):  # Itanium mangling: <length><identifier>
            if end - n >= 2 and name[end - n - 2:end - n].isdigit() and int(name[end - n - 2:end - n]) == n:
                base = name[end - n:end]
                break
            if end - n >= 1 and name[end - n - 1:end - n].isdigit() and int(name[end - n - 1:end - n]) == n:
                base = name[end - n:end]
                break
    if "ILb1" in name:
        base += "<true>"
    if "ILb0" in name:
        base += "<false>"
    m = re.search(r"ILi(\d+)E", name)
    if m:
        base += f"<{m.group(1)}>"
    return base


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_n"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    print("# static summary of echoglad_b200/libechoglad_b200.so (sm_100a): instructions, resource usage, key mnemonics")
    for fn, c in sorted(counts.items(), key=lambda kv: -kv[1]["_n"]):
        print(f"{short(fn)}: {c['_n']} SASS instructions; {usage.get(fn, '')}")
        print("    " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))


if __name__ == "__main__":
    main()
