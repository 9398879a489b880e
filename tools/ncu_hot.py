"""Per line range of one source file (outermost inline frame): static SASS size, hot SASS size (instructions executed
more than HOT times), share of executed warp instructions and of stall samples, and the no_instruction samples.
usage: ncu_hot.py <prof_src.csv> <lib.so> <kernel-substr> <file> name:lo-hi ..."""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, so, kernel, fname = sys.argv[1:5]
HOT = float(os.environ.get("HOT", "1e5"))
ranges = []
for a in sys.argv[5:]:
    nm, r = a.split(":"); lo, hi = r.split("-"); ranges.append((nm, int(lo), int(hi)))
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
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln):
        lines.append(chain[-1] if chain else None)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(body) == len(lines), (len(body), len(lines))
f = lambda r, k: float(r[ix[k]] or 0)
tot_i = sum(f(r, "Instructions Executed") for r in body); tot_s = sum(f(r, "# Samples") for r in body)
noi_key = "stall_no_inst" if "stall_no_inst" in ix else None
print(f"kernel: {len(body)} SASS instructions ({len(body) * 16 / 1024:.1f} KB), {tot_i:.4g} executed, {tot_s:.0f} samples")
for nm, lo, hi in ranges:
    st = hot = 0; i = s = t = ni = 0.0
    for r, key in zip(body, lines):
        if key and key[0] == fname and lo <= key[1] <= hi:
            st += 1; e = f(r, "Instructions Executed"); hot += e > HOT
            i += e; s += f(r, "# Samples"); t += f(r, "Thread Instructions Executed")
            if noi_key: ni += f(r, noi_key)
    print(f"{nm:14s} L{lo}-{hi}: static {st:5d} hot {hot:5d} ({hot * 16 / 1024:5.1f} KB) | {i / tot_i * 100:5.1f}% inst {s / tot_s * 100:5.1f}% smp"
          f" lanes {t / max(i, 1):4.1f} no_inst {ni / max(s, 1) * 100:4.1f}% of its smp")
