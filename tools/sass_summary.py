#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: which Blackwell-specific instructions each kernel of liblzy_b200.so really
contains (UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDG/STG widths and cache operators, SHFL, RED/ATOM, MEMBAR, CCTL).
Runs here, without a GPU:  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

SO = Path(__file__).resolve().parent.parent / "lzy_b200" / "lib" / "liblzy_b200.so"
INTEREST = re.compile(r"\b(UBLKCP[.\w]*|UTMA\w+[.\w]*|SYNCS[.\w]*|LDG[.\w]*|STG[.\w]*|LDS[.\w]*|STS[.\w]*|SHFL[.\w]*|REDG?[.\w]*|ATOMG?[.\w]*|"
                      r"MEMBAR[.\w]*|CCTL[.\w]*|FENCE[.\w]*|NANOSLEEP|ERRBAR|BAR[.\w]*|LDGSTS[.\w]*|UTC\w+|LDTM|STTM|HMMA[.\w]*)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(SO)], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# {SO.name}: cubin architectures {arch}")
    print("# counts of the instructions that tell the access path; full listing: cuobjdump -sass lzy_b200/lib/liblzy_b200.so\n")
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        if name:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", line)
            if m and INTEREST.match(m.group(1)):
                kernels[name][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    for (mangled, counts), pretty in zip(kernels.items(), demangle):
        print(f"## {pretty}")
        if mangled in usage:
            print(f"   {usage[mangled]}")
        for op, n in sorted(counts.items()):
            print(f"   {n:5d}  {op}")
        print()


if __name__ == "__main__":
    sys.exit(main())
