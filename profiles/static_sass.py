"""Static evidence from the built library (no GPU needed): per-kernel registers / spills and the SASS mnemonics that
show which hardware path a kernel uses (multimem = LDGMC/STGMC... , TMA bulk copy = UBLKCP, 128-bit LDG/STG, FFMA).

    python profiles/static_sass.py > profiles/r01_static_sass.txt
"""
import collections
import re
import subprocess
import sys

SO = "accelerate_b200/lib/libb200dp.so"
INTEREST = ["LDGMC", "STGMC", "REDGMC", "UBLKCP", "SYNCS", "LDG.E.128", "STG.E.128", "LDG.E.EF.128", "STG.E.EF.128",
            "FFMA", "MUFU.RSQ", "MUFU.RCP", "LD.E.STRONG.SYS", "ST.E.STRONG.SYS", "MEMBAR", "ERRBAR", "CCTL"]


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        elif fn and "REG:" in line:
            usage[fn] = dict(kv.split(":") for kv in line.split() if ":" in kv)
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for key in INTEREST:
                if op.startswith(key) or key in op:
                    counts[fn][key] += 1
            if ".128" in op and op.split(".")[0] in ("LDG", "STG", "LD", "ST") and op not in INTEREST:
                counts[fn][op] += 1
    print(f"# {SO}: sm_100a SASS, per kernel: registers, spill bytes (LOCAL), static smem, selected mnemonics")
    for fn in sorted(usage, key=demangle):
        u = usage[fn]
        picks = ", ".join(f"{k} x{v}" for k, v in sorted(counts[fn].items()))
        print(f"{demangle(fn)}\n    REG {u.get('REG')}  LOCAL {u.get('LOCAL')}  STACK {u.get('STACK')}  SHARED {u.get('SHARED')}  | {picks}")
    if any(int(u.get("LOCAL", 0)) for u in usage.values()):
        print("# WARNING: a kernel spills", file=sys.stderr)


if __name__ == "__main__":
    main()
