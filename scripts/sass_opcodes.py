"""SASS evidence: per-kernel counts of the tcgen05 / TMA / TMEM opcodes in the in-tree library.
Usage: python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt   (cuobjdump -sass on liblwpose_b200.so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lightweight-human-pose-estimation.pytorch_b200", "liblwpose_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "FFMA2", "HMNMX2",
        "F2FP", "MUFU", "DFMA", "DADD", "DMUL", "ELECT", "ACQBULK", "FENCE"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    full = op if k in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UBLKCP") else k
                    cur[full] += 1
    print("# cuobjdump -sass %s : opcode counts per kernel (only kernels with tensor-core / TMA / TMEM opcodes or > 500 instructions)" % os.path.relpath(LIB, ROOT))
    tot = collections.Counter()
    for name, c in kernels.items():
        keys = {k: v for k, v in c.items() if k != "_total"}
        tot.update(keys)
        if not any(k.startswith(("UTC", "UTMA", "LDTM", "STTM", "UBLKCP")) for k in keys) and c["_total"] < 500:
            continue
        print("%s  [%d instructions]" % (name, c["_total"]))
        print("    " + ", ".join("%s=%d" % kv for kv in sorted(keys.items())))
    print("# library totals: " + ", ".join("%s=%d" % kv for kv in sorted(tot.items())))


if __name__ == "__main__":
    main()
