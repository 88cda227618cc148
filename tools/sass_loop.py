"""Static SASS view of a kernel's hot loop: finds the backward branch whose body is densest in
fp64 instructions (>= 15 of them) and prints the opcode histogram of that body.
    python tools/sass_loop.py curvis_b200/csrc/build/render_f64_fast.o FastEllisELi1 [--dump | --all]
"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
dump = "--dump" in sys.argv
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
body = next(f for f in funcs[1:] if pat in f.split("\n", 1)[0])
ins = []   # (addr, opcode, text)
for line in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        text = m.group(2).strip()
        t = re.sub(r"^@!?U?P\d+\s+", "", text)
        ins.append((int(m.group(1), 16), t.split()[0].split(".")[0], text))
addr_index = {a: i for i, (a, _, _) in enumerate(ins)}
best = None
for i, (a, op, text) in enumerate(ins):
    if op == "BRA":
        m = re.search(r"0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
            j = addr_index[int(m.group(1), 16)]
            n64 = sum(1 for k in range(j, i + 1) if ins[k][1] in ("DFMA", "DMUL", "DADD", "DSETP"))
            if n64 >= 15 and (best is None or n64 / (i - j + 1) > best[0] / (best[2] - best[1] + 1)):
                best = (n64, j, i)
if "--all" in sys.argv:   # every fp64-dense loop (an inlined helper may appear several times)
    for i, (a, op, text) in enumerate(ins):
        if op == "BRA":
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
                j = addr_index[int(m.group(1), 16)]
                n64 = sum(1 for k in range(j, i + 1) if ins[k][1] in ("DFMA", "DMUL", "DADD", "DSETP"))
                if n64 >= 15 and i - j + 1 < 200:
                    hist = collections.Counter(ins[k][1] for k in range(j, i + 1))
                    print(f"loop {ins[j][0]:#x}..{a:#x}: {i - j + 1} instructions, fp64-pipe {n64}: " + ", ".join(f"{o} {n}" for o, n in hist.most_common()))
    sys.exit(0)
n64, j, i = best
hist = collections.Counter(ins[k][1] for k in range(j, i + 1))
print(f"loop {ins[j][0]:#x}..{ins[i][0]:#x}: {i - j + 1} instructions, fp64-pipe {n64}")
for op, n in hist.most_common():
    print(f"  {op:10s} {n}")
if dump:
    for k in range(j, i + 1):
        print(f"{ins[k][0]:#06x}  {ins[k][2]}")
