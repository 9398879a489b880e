"""Sum ncu source-page instruction / sample shares over line ranges of one source file (outermost
inline frame).  usage: ncu_ranges.py <prof_src.csv> <lib.so> <kernel-substr> <file> name:lo-hi ..."""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, so, kernel, fname = sys.argv[1:5]
ranges = []
for a in sys.argv[5:]:
    nm, r = a.split(":"); lo, hi = r.split("-"); ranges.append((nm, int(lo), int(hi)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, chain, inside, ops = [], [], False, []
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
    mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if mm:
        lines.append(chain[-1] if chain else None)
        ops.append(mm.group(1))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(body) == len(lines), (len(body), len(lines))
tot_i = sum(float(r[ix["Instructions Executed"]] or 0) for r in body)
tot_s = sum(float(r[ix["# Samples"]] or 0) for r in body)
for nm, lo, hi in ranges:
    i = s = t = 0.0; pipes = collections.Counter()
    for r, key, op in zip(body, lines, ops):
        if key and key[0] == fname and lo <= key[1] <= hi:
            e = float(r[ix["Instructions Executed"]] or 0)
            i += e; s += float(r[ix["# Samples"]] or 0); t += float(r[ix["Thread Instructions Executed"]] or 0)
            o = op.split()[0] if not op.startswith("@") else op.split()[1]
            pipes[o.split(".")[0]] += e
    top = ", ".join(f"{k} {v/max(i,1)*100:.0f}%" for k, v in pipes.most_common(8))
    print(f"{nm:10s} L{lo}-{hi}: {i/tot_i*100:5.1f}% inst ({i:.3g}) {s/tot_s*100:5.1f}% smp lanes {t/max(i,1):4.1f} | {top}")
