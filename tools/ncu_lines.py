"""Per-line (outermost frame) instruction counts per block for a line range. usage: ncu_lines.py csv so kernel file lo hi nblocks"""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, so, kernel, fname = sys.argv[1:5]
lo, hi, nb = int(sys.argv[5]), int(sys.argv[6]), float(sys.argv[7])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, chain, inside = [], [], False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kernel in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        f, l = os.path.basename(m.group(1)), int(m.group(2))
        if m.group(3) is None: chain = [(f, l)]
        else:
            if not chain or chain[-1] != (f, l): chain = [(f, l)]
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(chain[-1] if chain else None)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(body) == len(lines)
inst, smp, n_sass = collections.Counter(), collections.Counter(), collections.Counter()
for r, key in zip(body, lines):
    if key and key[0] == fname and lo <= key[1] <= hi:
        inst[key[1]] += float(r[ix["Instructions Executed"]] or 0); smp[key[1]] += float(r[ix["# Samples"]] or 0); n_sass[key[1]] += 1
src = open("integrated_path_planning_b200/csrc/" + fname).read().splitlines()
ts = sum(float(r[ix["# Samples"]] or 0) for r in body)
for l in sorted(inst):
    if inst[l] / nb < 20 and smp[l]/ts < 0.003: continue
    print(f"L{l:4d} {inst[l]/nb:8.0f} inst/blk {smp[l]/ts*100:5.1f}% smp {n_sass[l]:4d} sass | {src[l-1].strip()[:100]}")
