#!/usr/bin/env python
"""Aggregate an ncu --set full --import-source report by CUDA source line.

usage: ncu_lines.py report.ncu-rep libpqdet_b200.so kernel_mangled_substring [top_n]
Maps SASS addresses to source lines through `nvdisasm -g` of the cubin embedded in the .so.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    dis = None
    for f in sorted(os.listdir(tmp)):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in out and ".text." in out:
            dis = out
    lines = dis.split("\n")
    start = [i for i, l in enumerate(lines) if l.strip().startswith(".section") and ".text." in l and kname in l][0]
    addr2line, cur = {}, None
    for l in lines[start + 1:]:
        if l.strip().startswith(".section") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            addr2line[int(m.group(1), 16)] = cur
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = {c: hdr.index(c) for c in hdr if c.startswith("stall_") and "Not Issued" not in c}
    agg, samp, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    base, n = None, 0
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
            if n and r and r[0] == "Kernel Name":
                break
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        k = addr2line.get(a - base)
        agg[k] += int(r[ii]); samp[k] += int(r[isamp]); n += 1
        for c, i in stall.items():
            if r[i] and int(r[i]):
                st[k][c] += int(r[i])
    tot, ts = sum(agg.values()), sum(samp.values())
    print("total warp instructions %d, stall samples %d" % (tot, ts))
    allst = collections.Counter()
    for k in st:
        allst.update(st[k])
    print("stall mix:", ", ".join("%s %.1f%%" % (c.replace("stall_", ""), 100.0 * v / ts) for c, v in allst.most_common(8)))
    for k, v in sorted(agg.items(), key=lambda x: (-samp[x[0]] if os.environ.get("SORT", "samples") == "samples" else -x[1]))[:top]:
        tops = ", ".join("%s:%d" % (c.replace("stall_", ""), x) for c, x in st[k].most_common(3))
        print("%-34s inst %5.1f%%  samples %5.1f%%  %s" % (k, 100.0 * v / tot, 100.0 * samp[k] / ts, tops))


if __name__ == "__main__":
    main()
