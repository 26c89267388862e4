"""Static SASS instruction counts per kernel of libpmcts_b200.so (a quick proxy before a GPU run).

  python scripts/sass_count.py [substring]

For every kernel: total instructions, and the opcode mix of its largest loop (the span of the backward
branch that encloses the most instructions).
"""
import re
import subprocess
import sys

import os
so = os.environ.get("SO", "iboat_pmcts_b200/libpmcts_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
funcs, name, ins = [], None, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            funcs.append((name, ins))
        name, ins = m.group(1), []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
    if m and name:
        ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
if name:
    funcs.append((name, ins))
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for name, ins in funcs:
    if flt not in name:
        continue
    dem = subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()
    best = (0, 0, 0)
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < addr:
                tgt = int(t.group(1), 16)
                if addr - tgt > best[0]:
                    best = (addr - tgt, tgt, addr)
    body = [(a, o) for a, o, _ in ins if best[1] <= a <= best[2]]
    ops = {}
    for _, o in body:
        o = o.split(".")[0]
        ops[o] = ops.get(o, 0) + 1
    print(len(ins), "loop", len(body), dem[:80])
    if flt:
        print("   ", sorted(ops.items(), key=lambda kv: -kv[1]))
