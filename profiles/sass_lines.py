#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples of one kernel: joins the SASS page of an .ncu-rep (executed
instructions and warp-stall samples per SASS address) with nvdisasm -g line info of the in-tree library.
usage: sass_lines.py file.ncu-rep <mangled-name-substring> [top_n]"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("GI_LIB") or os.path.join(ROOT, "gi_raytracer_b200", "libgi_b200.so")   # the library the profiled run loaded


def line_map(sym):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, stdout=subprocess.DEVNULL, check=True)
    cub = max((os.path.join(d, f) for f in os.listdir(d)), key=os.path.getsize)
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
    m, cur, inside = {}, None, False
    for l in txt:
        if l.startswith("\t.section\t.text."):
            inside = sym in l
            continue
        if not inside:
            continue
        g = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if g:
            cur = (os.path.basename(g.group(1)), int(g.group(2)))
            continue
        g = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if g:
            m[int(g.group(1), 16)] = (cur, g.group(2).strip())
    return m


def main(rep, sym, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(raw))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    it = hdr.index("Thread Instructions Executed")
    lm = line_map(sym)
    body = [r for r in rows[hdr_i + 1:] if len(r) > isamp and r[ia].startswith("0x")]
    base = int(body[0][ia], 16)
    per_line, per_op = defaultdict(lambda: [0, 0, 0]), defaultdict(int)
    tot_i = tot_s = 0
    for r in body:
        off = int(r[ia], 16) - base
        n, s = int(r[ii] or 0), int(r[isamp] or 0)
        loc, ins = lm.get(off, (("?", 0), r[1].strip()))
        per_line[loc][0] += n
        per_line[loc][1] += s
        per_line[loc][2] += int(r[it] or 0)
        per_op[ins.split()[0].lstrip("@!P0123456789 ") if not ins.startswith("@") else ins.split()[1]] += n
        tot_i += n
        tot_s += s
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    print("--- by source line (warp instructions, share, stall-sample share, active lanes per instruction)")
    src_cache = {}
    for loc, (n, s, tn) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc and loc[0] != "?":
            p = os.path.join(ROOT, "gi_raytracer_b200", "csrc", loc[0])
            if p not in src_cache and os.path.exists(p):
                src_cache[p] = open(p).read().splitlines()
            if p in src_cache and 0 < loc[1] <= len(src_cache[p]):
                text = src_cache[p][loc[1] - 1].strip()[:90]
        print(f"{loc[0] if loc else '?'}:{loc[1] if loc else 0:<5d} {n:>14d} {100 * n / tot_i:5.1f}% {100 * s / max(tot_s, 1):5.1f}% {tn / max(n, 1):5.1f}  {text}")
    print("--- by opcode")
    for op, n in sorted(per_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{op:24s} {n:>14d} {100 * n / tot_i:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
