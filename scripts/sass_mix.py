#!/usr/bin/env python
"""Static opcode histogram of one kernel in an object/cubin: python scripts/sass_mix.py <obj> <name regex>"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], re.compile(sys.argv[2])
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, kernels = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur].append(m.group(3))
for name, ops in kernels.items():
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    if not pat.search(dem):
        continue
    c = collections.Counter(o.split(".")[0] for o in ops)
    print(dem[:150]); print("  total", len(ops), " ", "  ".join(f"{k}:{v}" for k, v in c.most_common(22)))
