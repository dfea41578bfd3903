"""Opcode evidence per hot kernel, from the built libtpdcu.so (cuobjdump -sass; no GPU needed): counts of the memory, atomic,
TMA/mbarrier and vote opcodes that the DESIGN.md claims rest on, plus registers / shared memory per kernel.
    python profiles/sass_evidence.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torpedo_b200", "lib", "libtpdcu.so")
KEEP = re.compile(r"^(LDG|STG|LDS|STS|ATOMS|ATOMG|RED|UBLKCP|UBLKPF|UTMA|SYNCS|VOTE|MATCH|SHFL|BAR|MUFU|LDGSTS|MEMBAR|FENCE|CCTL|R2P|POPC|HMMA|UTC|TCGEN)")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s+(.*)", res):
    usage[m.group(1)] = m.group(2).strip()
cur, ops, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        total[cur] += 1
        if KEEP.match(m.group(1)):
            ops[cur][m.group(1)] += 1
demangled = dict(zip(total, subprocess.run(["c++filt"] + list(total), capture_output=True, text=True).stdout.splitlines()))
print(f"# {os.path.relpath(LIB)}: sm_100a SASS opcode counts per kernel (static instruction counts, not executed counts)")
for k in sorted(total, key=lambda f: demangled[f]):
    name = re.sub(r"\(.*", "", demangled[k])
    print(f"\n{name}   [{total[k]} instructions; {usage.get(k, '')}]")
    row = sorted(ops[k].items(), key=lambda kv: (-kv[1], kv[0]))
    for i in range(0, len(row), 6):
        print("   " + "  ".join(f"{op} x{n}" for op, n in row[i:i + 6]))
