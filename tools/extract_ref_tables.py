#!/usr/bin/env python3
"""Parse the literal constant tables of the reference (GEMMul8/src/table.hpp) into
tests/golden/ref_tables.json.  Run in the build container (needs /root/reference); the JSON is
committed so that tests can pin our *derived* constants (gemmul8_b200/tables.py) without the
reference tree being present.

    python tools/extract_ref_tables.py [/root/reference/GEMMul8/src/table.hpp]
"""
import json
import re
import sys
from pathlib import Path

HEXF = r"-?0x[01]\.[0-9a-fA-F]+p[+-]?\d+"


def block(src: str, ns: str, decl: str) -> str:
    """Text of `decl ... = { ... };` inside `namespace ns { ... }` (first match after the namespace opener)."""
    pat = re.compile(r"namespace\s+" + ns + r"\s*\{")
    for m in pat.finditer(src):
        d = src.find(decl, m.end())
        nxt = pat.search(src, m.end())
        other = re.compile(r"namespace\s+(INT8|FP8)\s*\{").search(src, m.end())
        if d < 0:
            continue
        if other and other.start() < d:
            continue
        start = src.index("{", src.index("=", d))
        depth, i = 0, start
        while True:
            c = src[i]
            depth += c == "{"
            depth -= c == "}"
            if depth == 0:
                return src[start:i + 1]
            i += 1
    raise KeyError((ns, decl))


def floats(txt):
    return [float.fromhex(t) for t in re.findall(HEXF, txt)]


def rows(txt):
    """Split a 2-level brace initialiser into rows."""
    inner = txt.strip()[1:-1]
    out, depth, cur = [], 0, ""
    for c in inner:
        if c == "{":
            depth += 1
            if depth == 1:
                cur = ""
                continue
        if c == "}":
            depth -= 1
            if depth == 0:
                out.append(cur)
                continue
        if depth >= 1:
            cur += c
    return out


def main():
    path = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/GEMMul8/src/table.hpp")
    src = path.read_text()
    out = {"source": str(path)}
    for be in ("INT8", "FP8"):
        d = {}
        d["moduli"] = [int(v) for _, v in sorted(
            ((int(i), v) for i, v in re.findall(r"moduli<gemmul8::Backend::" + be + r",\s*(\d+)>\s*=\s*(\d+)", src)))]
        d["log2P"] = {n: float.fromhex(v.rstrip("F")).hex() for n, v in
                      re.findall(r"log2P<gemmul8::Backend::" + be + r",\s*(\d+)>\s*=\s*(" + HEXF + r")F", src)}
        P = floats(block(src, be, "constexpr double2 P[19]"))
        d["P"] = [[P[2 * i].hex(), P[2 * i + 1].hex()] for i in range(19)]
        d["invP"] = [v.hex() for v in floats(block(src, be, "constexpr double invP[19]"))]
        d["qPi_1"] = [[v.hex() for v in floats(r)] for r in rows(block(src, be, "inline constexpr double qPi_1[19][20]"))]
        q2 = block(src, be, "inline constexpr double2 qPi_2[")
        d["qPi_2"] = []
        for r in rows(q2):
            f = floats(r)
            d["qPi_2"].append([[f[2 * i].hex(), f[2 * i + 1].hex()] for i in range(len(f) // 2)])
        mp = block(src, be, "mod_pow2_h[")
        d["mod_pow2"] = [[int(x) for x in re.findall(r"-?\d+", r)] for r in rows(mp)]
        out[be] = d
    out["sqrt_moduli"] = [int(v) for _, v in sorted(
        (int(i), v) for i, v in re.findall(r"sqrt_moduli<(\d+)>\s*=\s*(\d+)", src))]
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "ref_tables.json"
    dst.write_text(json.dumps(out, indent=1))
    print("wrote", dst, {be: (len(out[be]["moduli"]), len(out[be]["qPi_1"]), len(out[be]["qPi_2"])) for be in ("INT8", "FP8")})


if __name__ == "__main__":
    main()
