#!/usr/bin/env python
"""Mnemonic counts per kernel of the shipped libsfvos.so (`cuobjdump -sass`): the proof that the contraction kernels are
tcgen05 / TMEM / TMA.  usage: python tools/sass_summary.py > profiles/sass_<round>.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "applying-slowfast-networks-to-video-object-segmentation_b200", "libsfvos.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
rows, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); rows[cur] = dict.fromkeys(("UTCHMMA", "2CTA", "UTMALDG", "LDTM", "STTM", "DFMA", "HMMA"), 0)
        continue
    if cur is None:
        continue
    if "UTCHMMA" in line:
        rows[cur]["UTCHMMA"] += 1
        if ".2CTA" in line: rows[cur]["2CTA"] += 1
    elif re.search(r"\bHMMA\b", line): rows[cur]["HMMA"] += 1
    for k in ("UTMALDG", "LDTM", "STTM", "DFMA"):
        if re.search(r"\b" + k, line): rows[cur][k] += 1
print("# SASS summary of libsfvos.so (round 2, final code): `cuobjdump -sass libsfvos.so`, mnemonic counts per kernel (`tools/sass_summary.py`)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st (TMEM), DFMA = the fp64 validation-mode kernels\n")
print("| kernel (mangled) | UTCHMMA | of which .2CTA | UTMALDG | LDTM | STTM | DFMA |\n|---|---|---|---|---|---|---|")
tot = dict.fromkeys(("UTCHMMA", "2CTA", "UTMALDG", "LDTM", "STTM", "DFMA", "HMMA"), 0)
for k, v in rows.items():
    for a in tot: tot[a] += v[a]
    if any(v[a] for a in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "DFMA")):
        print(f"| `{k[:90]}` | {v['UTCHMMA']} | {v['2CTA']} | {v['UTMALDG']} | {v['LDTM']} | {v['STTM']} | {v['DFMA']} |")
print(f"| **total** | {tot['UTCHMMA']} | {tot['2CTA']} | {tot['UTMALDG']} | {tot['LDTM']} | {tot['STTM']} | {tot['DFMA']} |")
print(f"\n`HMMA` (mma.sync) instructions in the library: {tot['HMMA']}; wgmma does not exist on sm_100a.")
