#!/usr/bin/env python
"""Count the Blackwell-native SASS mnemonics per kernel of libsedt_b200.so (cuobjdump -sass): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA, HMMA = legacy mma.sync (must be absent).
    python tools/sass_evidence.py > profiles/r1k_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "sound_event_detection_transformer_b200", "lib", "libsedt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC\w*MMA\w*|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTMAPF|HMMA|HGMMA|UTCBAR|SYNCS|REDG|ATOMG)\b")
kern = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("sedt::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("__nv_bfloat16", "bf16")
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    if kern:
        for mm in pat.findall(line):
            key = "UTC*MMA" if mm.startswith("UTC") and "MMA" in mm else mm
            counts[kern][key] += 1
cols = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA"]
print(f"{'kernel':70s} " + " ".join(f"{c:>8s}" for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    if any(c[x] for x in cols):
        print(f"{k[:70]:70s} " + " ".join(f"{c[x]:8d}" for x in cols))
    tot.update(c)
print(f"{'TOTAL over ' + str(len(counts)) + ' kernels':70s} " + " ".join(f"{tot[x]:8d}" for x in cols))
