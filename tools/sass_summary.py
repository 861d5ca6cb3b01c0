"""Per-kernel instruction counts and the mnemonics that show the design (tcgen05 / TMA / mbarrier / 16-byte accesses) from
`cuobjdump -sass` of the built library.  Usage: python tools/sass_summary.py [path/to/libmixstage_b200.so] > profiles/...txt"""
import collections
import os
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "RED", "ATOM", "SYNCS", "LDG.E.128", "STG.E.128", "LD.E.128", "ST.E.128",
        "LDS.128", "STS.128")


def main(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fn, per = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            per[fn][m.group(1)] += 1
    names = list(per)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    rows = []
    for f, name in zip(names, dem):
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
        sel = collections.Counter()
        for k, n in per[f].items():
            for key in KEYS:
                if k.startswith(key):
                    sel[key] += n
        rows.append((short, sum(per[f].values()), " ".join("%s=%d" % kv for kv in sorted(sel.items()))))
    for short, tot, sel in sorted(rows):
        print("%-72s %6d  %s" % (short[:72], tot, sel))


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "mixstage_b200", "libmixstage_b200.so"))
