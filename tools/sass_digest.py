#!/usr/bin/env python
"""Counts of the Blackwell-native instructions in the shipped library (cuobjdump -sass), whole library and per kernel.
    python tools/sass_digest.py > profiles/rNN/sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "kindergarten-vq-vae_b200", "libkvq.so")
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG.2D.2CTA", "UTMALDG", "UTMAPF", "UTCBAR", "UBLKCP", "SYNCS",
            "FMNMX3.NAN", "ELECT", "REDUX|MATCH.ANY", r"RED\.|REDG", r"ATOM\.|ATOMG|ATOMS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    total = collections.Counter()
    per_fn = collections.OrderedDict()
    size = collections.Counter()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per_fn[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
        if not m or fn is None:
            continue
        size[fn] += 1
        ins = m.group(1)
        for p in PATTERNS:
            if re.search(p.replace(".2CTA", r"\.2CTA").replace(".2D", r"\.2D").replace(".NAN", r"\.NAN") if "|" not in p and "\\" not in p else p, ins):
                total[p] += 1
                per_fn[fn][p] += 1
    print(f"# SASS digest of {os.path.relpath(SO, ROOT)} (cuobjdump -sass; tools/sass_digest.py)")
    print("# occurrences of the Blackwell-native instructions in the whole library")
    for p in PATTERNS:
        print(f"{p}: {total[p]}")
    print("\n# per kernel: SASS instructions | UTCHMMA / LDTM / UTMALDG / UTMAPF / UBLKCP")
    for fn, c in sorted(per_fn.items()):
        if c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"] or c["UBLKCP"]:
            print(f"{fn} : {size[fn]} | {c['UTCHMMA']} {c['LDTM']} {c['UTMALDG']} {c['UTMAPF']} {c['UBLKCP']}")


if __name__ == "__main__":
    sys.exit(main())
