#!/usr/bin/env python
"""SASS opcode histogram per kernel of the built library (cuobjdump -sass), for profiles/rNN_sass_hist.txt.

  python tools/sass_hist.py [--lib path.so] [--filter regex] [--top N]

Static counts (instructions in the binary, not executed counts); loops in the hot kernels are fully unrolled, so
for k_row32 / k_col_* the static count of the straight-line body is what one thread executes per row / tile.
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "audio_matcher_b200", "libaudio_matcher_b200.so"))
    ap.add_argument("--filter", default=r"k_row32<13, 0>|k_row32_stream<13, [03]>|k_col_fwd_stream<9, 4, 32, 1, 13>|k_col_inv<9, 4, 32, 13>|k_col_inv<9, 4, 32, 14>|k_row32<14, 0>|k_tile_from_runs|k_chunk_peaks<true>")
    ap.add_argument("--top", type=int, default=18)
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void amk::", "").replace("void amp::", "").replace("void (anonymous namespace)::", "")
            cur = name if re.search(args.filter, name) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op, mods = m.group(1), m.group(2)
            if op in ("LDG", "STG", "LDS", "STS", "LD", "ST", "UTMALDG", "UBLKCP", "SYNCS", "BAR"):
                op += "".join(x for x in mods.split(".") if x in ("64", "128", "E", "2D", "S", "G", "ARRIVE", "TRANS64", "SYNC"))and ("." + ".".join(x for x in mods.split(".") if x in ("64", "128", "2D", "ARRIVE", "TRANS64", "SYNC")) ) or ""
            hist[cur][op] += 1
    for name, h in hist.items():
        tot = sum(h.values())
        fp = sum(v for k, v in h.items() if k in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2"))
        print(f"{name}: {tot} instructions, FADD/FMUL/FFMA {fp} ({100.0 * fp / max(tot, 1):.1f} %)")
        print("   " + "  ".join(f"{k} {v}" for k, v in h.most_common(args.top)))


if __name__ == "__main__":
    main()
