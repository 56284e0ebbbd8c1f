#!/usr/bin/env python
"""Static SASS instruction counts per source line of one kernel of libfilo_b200.so.
usage: sass_lines.py <mangled-name-substring> [file-substring] [lo] [hi]"""
import collections, os, re, subprocess, sys, tempfile
kname = sys.argv[1]; fsub = sys.argv[2] if len(sys.argv) > 2 else ""; lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0; hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "filodb_b200", "libfilo_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cnt = collections.Counter(); tot = 0
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    fn = cur = None
    for ln in dis.splitlines():
        m = re.search(r'\.text\.(\S+):', ln)
        if m: fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        if fn and kname in fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln): cnt[cur] += 1; tot += 1
print("static instructions:", tot)
srcs = {}
for (f, l), v in sorted(cnt.items()):
    if fsub in f and lo <= l <= hi:
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "filodb_b200", "csrc", f)
            srcs[f] = open(p).read().split("\n") if os.path.exists(p) else None
        t = srcs[f][l - 1][:120] if srcs[f] and l <= len(srcs[f]) else ""
        print("%-20s %5d %5d  %s" % (f, l, v, t))
