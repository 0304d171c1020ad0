"""Developer tool: static opcode histogram of the stage loop (largest backward branch span below 40 KB) of one flight
kernel instance in a built library.  usage: python tools/sass_loop_hist.py <lib.so> [mangled-name-substring]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "emc_flight_kernelILi128ELi3ELi2ELi1ELi1"
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
blocks = txt.split("Function : ")
body = next(b for b in blocks if want in b.split("\n", 1)[0])
ins = []
for l in body.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = (0, 0)
for a, t in ins:
    if "BRA" in t:
        mm = re.search(r"0x([0-9a-f]+)", t)
        if mm:
            tgt = int(mm.group(1), 16)
            if tgt < a and 4000 < a - tgt < 40000 and a - tgt > best[1] - best[0]:
                best = (tgt, a)
ops = collections.Counter()
for a, t in ins:
    if best[0] <= a <= best[1]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
tot = sum(ops.values())
fp64 = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"{want}: total {len(ins)} instructions; stage loop {hex(best[0])}..{hex(best[1])}: {tot} static, FP64 pipe {fp64}")
print("  " + "  ".join(f"{k} {v}" for k, v in ops.most_common(24)))
