#!/usr/bin/env python
"""Developer tool: md5 of the trace kernel's hot loop (the innermost loop with the 84 FFMA2), instruction text only —
addresses and branch targets removed. Two builds with the same signature run the same schedule in the loop that holds
99 % of the frame time, whatever else changed in the kernel.
    tools/hotloop_sig.py [lib.so ...]"""
import hashlib
import re
import subprocess
import sys


def signature(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    m = re.search(r"Function : _ZN3rtx12trace_kernelILb0.*?(?=Function : |\Z)", sass, re.S)
    ins = [(int(a, 16), t.strip()) for a, t in re.findall(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", m.group(0))]
    best = None
    for a, t in ins:
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t)
            if m2 and int(m2.group(1), 16) < a:
                body = [x for x in ins if int(m2.group(1), 16) <= x[0] <= a]
                if sum(1 for x in body if x[1].startswith("FFMA2")) == 84 and (best is None or len(body) < len(best)):
                    best = body
    if best is None:
        return None, 0, len(ins)
    text = "\n".join(re.sub(r"0x[0-9a-f]+", "ADDR", t) if "BRA" in t else t for _, t in best)
    return hashlib.md5(text.encode()).hexdigest(), len(best), len(ins)


for lib in sys.argv[1:] or ["ray-tracer-from-scratch_b200/librtx_b200.so"]:
    sig, n, total = signature(lib)
    print("%s  hot loop %d instr  kernel %d instr  %s" % (sig, n, total, lib))
