#!/usr/bin/env python
"""Static SASS instruction count of one kernel by source line range (code-size / I-cache footprint audit).
usage: sass_size.py lib.so kernel_mangled_substring"""
import collections, os, re, subprocess, sys, tempfile
so, kname = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = None
for f in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname in out and ".text." in out:
        dis = out
lines = dis.split("\n")
start = [i for i, l in enumerate(lines) if l.strip().startswith(".section") and ".text." in l and kname in l][0]
cnt, cur, total = collections.Counter(), None, 0
for l in lines[start + 1:]:
    if l.strip().startswith(".section") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l):
        cnt[cur] += 1
        total += 1
print("total SASS instructions: %d (%.0f KB)" % (total, total * 16 / 1024.0))
byfile = collections.Counter()
for (f, n), c in cnt.items():
    byfile[f] += c
print("by file:", dict(byfile))
for (f, n), c in sorted(cnt.items(), key=lambda x: -x[1])[:40]:
    print("%-28s %5d  %5d" % (f, n, c))
