"""Like ncu_by_line.py but attributes every SASS instruction to the OUTERMOST fot_kernels.cuh
line of its inline chain (nvdisasm -gi), i.e. to the statement of the kernel body that caused it.
usage: ncu_by_region.py <prof_src.csv> <libfot.so> <kernel-substr> [first_line last_line]"""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, so, kernel = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# find kernel-body line range: lines of fot_kernels.cuh that appear as non-inlined locations
lines, chain, inside = [], [], False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kernel in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        f, l = os.path.basename(m.group(1)), int(m.group(2))
        if m.group(3) is None:
            chain = [(f, l)]
        else:
            # chain entries come innermost first; keep appending the outer frames
            if not chain or chain[-1] != (f, l):
                chain = [(f, l)]
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(chain[-1] if chain else None)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(body) == len(lines), (len(body), len(lines))
inst, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r, key in zip(body, lines):
    inst[key] += float(r[ix["Instructions Executed"]] or 0)
    smp[key] += float(r[ix["# Samples"]] or 0)
    thr[key] += float(r[ix["Thread Instructions Executed"]] or 0)
ti, ts = sum(inst.values()), sum(smp.values())
src = open("integrated_path_planning_b200/csrc/" + os.environ.get("REGION_FILE", "fot_kernels.cuh")).read().splitlines()
print(f"total warp instructions {ti:.4g}")
for key in sorted(k for k in inst if k and k[0] == os.environ.get("REGION_FILE", "fot_kernels.cuh")):
    if inst[key] / ti < 0.004 and smp[key] / ts < 0.004: continue
    print(f"{inst[key]/ti*100:5.1f}% inst {smp[key]/ts*100:5.1f}% smp lanes {thr[key]/max(inst[key],1):4.1f}  L{key[1]}: {src[key[1]-1].strip()[:95]}")
other = sum(v for k, v in inst.items() if not k or k[0] != os.environ.get("REGION_FILE", "fot_kernels.cuh"))
print(f"{other/ti*100:5.1f}% inst attributed elsewhere")
