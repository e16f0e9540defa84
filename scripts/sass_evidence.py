#!/usr/bin/env python
"""SASS evidence for profiles/: which Blackwell instructions each shipped kernel of libb200awq.so contains
(cuobjdump -sass; mnemonics per /opt/skills/guides/B200_PROFILING.md: UTCHMMA / UTCIMMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = bulk-async copy, SYNCS = mbarrier, IMMA / HMMA = mma.sync,
LDGSTS = cp.async, LDSM = ldmatrix) and an excerpt of the decode kernel's inner loop.

    python scripts/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "llm_awq_b200", "lib", "libb200awq.so")
MNEM = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "IMMA", "HMMA",
        "LDSM", "LDGSTS", "REDUX", "UCGABAR", "ACQBULK"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line)
print("library:", os.path.relpath(LIB, ROOT), " arch: sm_100a")
print("%-110s %6s  %s" % ("kernel", "instr", "  ".join(MNEM)))
for name, lines in funcs.items():
    text = "\n".join(lines)
    counts = [len(re.findall(r"\b%s[.\w]*" % m, text)) for m in MNEM]
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("b200awq::", "")
    print("%-110s %6d  %s" % (short[:110], len(lines), "  ".join("%*d" % (len(m), c) for m, c in zip(MNEM, counts))))
# inner loop of the one-token decode kernel (16 consumer warps, fp16): from the first LDSM to the loop branch
for name, lines in funcs.items():
    if "w4a16_decode_kernel<__half, 16, false, false>" in name:
        idx = [i for i, l in enumerate(lines) if "LDSM" in l]
        if idx:
            print("\ninner loop of", name.split("(")[0], "(one group of 16 channels x 128 input channels per iteration):")
            for l in lines[idx[0] - 4: idx[0] + 48]:
                print("   ", re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l).strip())
