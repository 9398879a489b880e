"""Static SASS size of one kernel, total and per line range of one source file (outermost inline frame).
usage: sass_size.py <lib.so> <kernel-substr> [<file> name:lo-hi ...]"""
import os, re, subprocess, sys, tempfile
so, kernel = sys.argv[1:3]
fname = sys.argv[3] if len(sys.argv) > 3 else None
ranges = []
for a in sys.argv[4:]:
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
print(f"{kernel}: {len(lines)} SASS instructions = {len(lines) * 16 / 1024:.1f} KB")
for nm, lo, hi in ranges:
    n = sum(1 for k in lines if k and k[0] == fname and lo <= k[1] <= hi)
    print(f"  {nm:12s} L{lo}-{hi}: {n:5d} ({n * 16 / 1024:5.1f} KB)")
