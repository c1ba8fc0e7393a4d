#!/usr/bin/env python3
"""Static evidence that the hot kernels use the Blackwell paths (no GPU needed).  For every kernel of build/obj/g8_gemm_i8.o: counts of
the SASS mnemonics of tcgen05.mma (UTCIMMA = kind::i8, UTCQMMA = kind::f8f6f4; .2CTA = cta_group::2), tcgen05.ld (LDTM), tcgen05.commit
(UTCBAR), TMA tensor loads / stores (UTMALDG / UTMASTG), plain bulk copies (UBLKCP) and cluster barriers (UCGABAR); then ptxas'
registers / spills / shared memory for every kernel of the library (build/obj/*.ptxas.log).
usage: tools/sass_evidence.py > profiles/<tag>_sass_mnemonics.txt      (after `make -C gemmul8_b200/csrc`)"""
import collections, glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPI = {0: "MOD_I8", 1: "RAW_I32", 2: "BOUND_MAX", 3: "MOD_I8_CPLX", 4: "BOUND_MAX_CPLX", 5: "F8_MOD", 6: "F8_BOUND", 7: "F8_RAW", 8: "MOD_I8_SCATTER",
       9: "RAW_I32_SCATTER", 10: "F8_BOUND_CPLX", 11: "F8_PROD"}
KEYS = ["UTCIMMA", "UTCIMMA.2CTA", "UTCQMMA", "UTCQMMA.2CTA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UCGABAR_ARV"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "build/obj/g8_gemm_i8.o")], capture_output=True, text=True).stdout
counts, fn = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"\b(UTCIMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|LDTM|UCGABAR_ARV)((?:\.[A-Za-z0-9_]+)*)", line)
    if m and fn:
        key = m.group(1) + (".2CTA" if ".2CTA" in m.group(2) and m.group(1) in ("UTCIMMA", "UTCQMMA") else "")
        counts[fn][key] += 1
dm = demangle(list(counts))
print("# g8_gemm_i8.o: SASS mnemonic counts per kernel  (EPI name, CTA group)")
print(f"{'kernel':44s} " + " ".join(f"{k:>12s}" for k in KEYS))
rows = []
for fn, c in counts.items():
    m = re.search(r"gemm_i8_tc_kernel<(\d+), (\d+)>", dm[fn])
    name = f"gemm_i8_tc_kernel<EPI_{EPI[int(m.group(1))]}, CG={m.group(2)}>" if m else dm[fn][:44]
    rows.append((int(m.group(1)) if m else 99, int(m.group(2)) if m else 0, f"{name:44s} " + " ".join(f"{c.get(k, 0):12d}" for k in KEYS)))
for _, _, r in sorted(rows):
    print(r)

print("\n# ptxas -v per kernel: registers, barriers, static shared memory, spills  (build/obj/*.ptxas.log)")
for f in sorted(glob.glob(os.path.join(ROOT, "build/obj/*.ptxas.log"))):
    txt, fn, spill = open(f).read().splitlines(), None, ""
    items = []
    for line in txt:
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            fn = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = f"spill {m.group(2)}/{m.group(3)} B"
        m = re.search(r"Used (\d+) registers(.*)", line)
        if m and fn:
            items.append((fn, f"{m.group(1)} regs{m.group(2)}; {spill}"))
    dm = demangle([i[0] for i in items])
    print(f"## {os.path.basename(f).replace('.ptxas.log', '.cu')}")
    for fn, info in items:
        print(f"  {re.sub(r'\(.*', '', dm[fn])[:90]:90s} {info}")
