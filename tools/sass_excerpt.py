"""SASS evidence for profiles/: per kernel of libfot.so, the instruction count and the counts of the mnemonics the
design relies on.  usage: python tools/sass_excerpt.py > profiles/rN/sass_excerpt.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "integrated_path_planning_b200", "libfot.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
print("SASS evidence: `cuobjdump -sass integrated_path_planning_b200/libfot.so` (sm_100a, nvcc 12.9), built from this commit's sources.")
print("Per kernel: instruction count and the counts of the mnemonics the design relies on.")
print("  UBLKCP = cp.async.bulk global->shared (TMA bulk copy), SYNCS = mbarrier ops, LDGSTS = cp.async, REDUX/CREDUX = warp redux,")
print("  MATCH = match.any, VOTE = ballot/any/all, DFMA/DMUL/DADD/DSETP = FP64 pipe, BAR = block barrier, ATOMS = shared atomics,")
print("  NANOSLEEP = back-off of the gated launch.  No UTMALDG / UTCHMMA / HMMA / tcgen05 anywhere: the path has no contraction")
print("  (K <= 6), tensor cores are unused by design.\n")
keys = ["UBLKCP", "SYNCS", "LDGSTS", "REDUX", "CREDUX", "MATCH", "VOTE", "DFMA", "DMUL", "DADD", "DSETP", "BAR", "ATOMS", "NANOSLEEP",
        "UTMALDG", "UTCHMMA", "HMMA", "MUFU"]
first = collections.OrderedDict()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", f, flags=re.M)
    cnt = collections.Counter()
    for i in ins:
        body = i.split()
        op = body[1] if body[0].startswith("@") and len(body) > 1 else body[0]
        base = op.split(".")[0]
        cnt[base] += 1
        if base in ("UBLKCP", "SYNCS", "REDUX", "CREDUX", "MATCH", "NANOSLEEP", "LDGSTS", "BAR") and (base, op) not in first:
            first[(base, op)] = (name, i)
    print(f"{name}\n    {len(ins)} instructions; " + ", ".join(f"{k} {cnt[k]}" for k in keys if cnt[k]))
print("\nfirst occurrence of each variant of the Blackwell-path mnemonics:")
for (base, op), (n, i) in first.items():
    print(f"  {op:24s} {i:60s} in {n[:48]}")
