#!/usr/bin/env python
"""Pair-loop lengths of step_kernel_c in a built library: for every backward branch of the <8, 1> instance whose body
holds MUFU instructions, the number of instructions of the body and its MUFU / LDS / LDC counts.

    python tools/sass_loops.py [library] [kernel-substring]
"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "particle_simulator_b200/libpsim_b200.so"
want = sys.argv[2] if len(sys.argv) > 2 else "step_kernel_cILi8ELi1"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, fn = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        fn[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        fn[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in fn.items():
    if want not in name:
        continue
    addr = {a: k for k, (a, _) in enumerate(ins)}
    loops = []
    for k, (a, text) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)* (?:\w+, )?`?\(?0x([0-9a-f]+)", text) or re.search(r"BRA.* 0x([0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt in addr and addr[tgt] <= k:
            body = [t for _, t in ins[addr[tgt]:k + 1]]
            mufu = sum("MUFU" in t for t in body)
            if mufu:
                loops.append((len(body), mufu, sum("LDS" in t for t in body), sum("LDC" in t for t in body)))
    print(f"{name}: {len(ins)} instructions; loops (instructions, MUFU, LDS, LDC): {loops}")
