"""Per-kernel SASS opcode histogram of libb200diar.so (evidence for profiles/: UTC*MMA = tcgen05.mma, UTMALDG = TMA,
LDTM/STTM = tcgen05.ld/st, UTCCP = tcgen05.cp, UTCBAR = tcgen05.commit).  Usage: python scripts/sass_histogram.py > out.md"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "pyannote_audio_b200/lib/libb200diar.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCCP", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "FFMA2", "FFMA",
        "DFMA", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "SYNCS", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(name.split("(")[0].replace("b200::", ""), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1).split(".")[0]
            cur[op] += 1
            cur["_total"] += 1
    print("| kernel | instructions | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for name, c in kernels.items():
        print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c.get(k, 0) or "") for k in KEYS) + " |")


if __name__ == "__main__":
    main()
