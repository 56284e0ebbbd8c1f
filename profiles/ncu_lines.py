#!/usr/bin/env python
"""Attribute executed SASS instructions of an ncu report to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <kernel-name-substring> [libfilo_b200.so]
Joins `ncu --page source --csv` (per-SASS-instruction counters) with `nvdisasm -g` line info of the matching cubin."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "filodb_b200", "libfilo_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn = cur = None
    for ln in dis.splitlines():
        m = re.search(r'\.text\.(\S+):', ln)
        if m: cur_fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
        if m and cur_fn and kname in cur_fn: addr2line.setdefault(cur_fn, {})[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kernel = rows[0][1]
hdr = rows[1]; ia, ie, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = int(rows[2][ia], 16)
# pick the function whose mangled name matches the demangled kernel best (same instruction count)
best = max(addr2line.items(), key=lambda kv: -abs(len(kv[1]) - (len(rows) - 2)))[1]
agg, samp, tot, tsamp = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    key = best.get(int(r[ia], 16) - base, ('?', 0))
    n, s = int(r[ie]), int(r[isamp]); agg[key] += n; samp[key] += s; tot += n; tsamp += s
print("kernel:", kernel[:120]); print("SASS instructions:", len(rows) - 2, " executed warp-instructions:", tot)
byfile = collections.Counter()
for (f, l), v in agg.items(): byfile[f] += v
print("by file:", {k: "%.1f%%" % (100 * v / tot) for k, v in byfile.most_common()})
for k, v in agg.most_common(int(os.environ.get("TOP", "40"))):
    print("%-22s %5d  %12d  %5.1f%%  stall-samples %5.1f%%" % (k[0], k[1], v, 100 * v / tot, 100 * samp[k] / max(tsamp, 1)))
