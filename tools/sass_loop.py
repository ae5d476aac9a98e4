#!/usr/bin/env python
"""Dump the SASS of one kernel of libbh_b200.so and count the instructions between consecutive
256-bit record loads of the walk's unrolled loop (= issue slots per visit).
    python tools/sass_loop.py k_walkILi2E [--dump out.sass]"""
import re
import subprocess
import sys

so = "barnes-hut-n-body_b200/csrc/libbh_b200.so"
pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", b)]
    print(name, len(ins), "instructions")
    if "--dump" in sys.argv:
        open(sys.argv[sys.argv.index("--dump") + 1], "w").write("\n".join(f"{a:05x} {t}" for a, t in ins) + "\n")
    loads = [i for i, (_, t) in enumerate(ins) if "LDG.E.ENL2.256" in t]
    gaps = [b - a for a, b in zip(loads, loads[1:])]
    print("record loads:", len(loads), "instructions between consecutive loads:", gaps)
    from collections import Counter
    if len(loads) > 3:
        seg = ins[loads[1]:loads[2]]
        c = Counter(t.split()[1] if t.startswith("@") else t.split()[0] for _, t in seg)
        print("one visit:", dict(c.most_common()))
