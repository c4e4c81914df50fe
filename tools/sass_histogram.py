"""Instruction histogram per kernel of libedb.so (evidence that the hot kernels are tcgen05 / TMEM /
TMA code and not a recompiled library): counts of the SASS mnemonics named in
/opt/skills/guides/B200_PROFILING.md per `Function :` section of `cuobjdump -sass`.

usage: python tools/sass_histogram.py [easydist_b200/libedb.so] > profiles/r02_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF",
        "SYNCS", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "RED", "ATOMG",
        "HMMA", "MUFU.TANH"]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "easydist_b200/libedb.so"
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    demangle = {}
    cur, hist, total = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[cur][k] += 1
                if k == "UTCHMMA" and ".2CTA" in op:
                    hist[cur]["UTCHMMA.2CTA"] += 1
    names = sorted(total, key=lambda f: -total[f])
    try:
        dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    print(f"# cuobjdump -sass {path}: {len(names)} kernels, {sum(total.values())} SASS instructions")
    print("# columns: instructions | " + " ".join(KEYS + ["UTCHMMA.2CTA"]))
    for f in names:
        h = hist[f]
        if not h:
            continue
        name = demangle.get(f, f)
        name = re.sub(r"\(.*", "", name)[:100]
        print(f"{name}: {total[f]} | " + " ".join(f"{k}={h[k]}" for k in KEYS + ["UTCHMMA.2CTA"] if h[k]))


if __name__ == "__main__":
    main()
