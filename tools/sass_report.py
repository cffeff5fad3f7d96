"""Static evidence, no GPU needed: per kernel of libvoicesplit_sm100.so the register / shared-memory usage (cuobjdump -res-usage)
and how many tcgen05 / TMEM / TMA instructions its SASS holds (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG / UTMASTG / UBLKCP = cp.async.bulk[.tensor], SYNCS = mbarrier) - B200_PROFILING.md's "which SASS proves what" table.

    python tools/sass_report.py > profiles/r01_sass_report.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "voicesplit_b200", "libvoicesplit_sm100.so")
PAT = {"tcgen05.mma": r"\bUTC[A-Z]*MMA", "tcgen05.ld/st": r"\b(LDTM|STTM)", "tma": r"\b(UTMALDG|UTMASTG|UBLKCP|UTMAPF)", "tcgen05.commit": r"\bUTCBAR",
       "mbarrier": r"\bSYNCS", "mufu": r"\bMUFU", "hfma/ffma": r"\b(FFMA|HFMA2)", "dfma": r"\bDFMA"}


def demangle(name):
    m = re.search(r"_ZN2vs\d+([A-Za-z0-9_]+?)I", name) or re.search(r"_ZN2vs\d+([A-Za-z0-9_]+?)E", name)
    short = m.group(1) if m else name
    targs = re.search(r"ILi(\d+)(?:ELi(\d+))?(?:ELi(\d+))?", name)
    return short + ("<" + ",".join(g for g in targs.groups() if g is not None) + ">" if targs else "")


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL)\s*:\s*(\d+)", line))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(lambda: collections.Counter())
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        for key, pat in PAT.items():
            if re.search(pat, line):
                counts[cur][key] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(usage)} kernels (sm_100a SASS); columns: registers, stack bytes, static smem, then SASS instruction counts")
    hdr = ["kernel", "REG", "STACK", "SHARED"] + list(PAT)
    print(" | ".join(hdr))
    for fn in sorted(usage, key=lambda f: (-counts[f]["tcgen05.mma"], demangle(f))):
        u = usage[fn]
        row = [demangle(fn), u.get("REG", "?"), u.get("STACK", "0"), u.get("SHARED", "0")] + [str(counts[fn][k]) for k in PAT]
        print(" | ".join(row))
    spills = [demangle(f) for f, u in usage.items() if int(u.get("STACK", "0")) > 0]
    print(f"# kernels with a stack frame (possible spills / local arrays): {spills if spills else 'none'}")


if __name__ == "__main__":
    sys.exit(main())
