"""Writes a copy of a .cu file for the cusim host build: `extern __shared__` stays an extern declaration (the harness defines
filo::smem) and function-scope `__shared__` variables become `static` (one CTA runs at a time, so a static is CTA-shared).
usage: make_cusim_src.py <in.cu> <out.cu>"""
import re
import sys

src = open(sys.argv[1]).read()
src = src.replace("extern __shared__", "extern /*shared*/")
src = re.sub(r"(?<![A-Za-z_])__shared__(?![A-Za-z_])", "static", src)
open(sys.argv[2], "w").write(src)
