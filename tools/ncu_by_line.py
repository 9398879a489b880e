"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line.

usage: ncu_by_line.py <prof_src.csv> <libfot.so> <kernel-mangled-substring> [top]
Maps the i-th SASS instruction of the kernel (nvdisasm -g line annotations, including the
inlined-at chain's innermost line) to ncu's i-th row and sums executed instructions / samples.
"""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, so, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kernel in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(body) == len(lines), (len(body), len(lines))
inst, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r, key in zip(body, lines):
    inst[key] += float(r[ix["Instructions Executed"]] or 0)
    smp[key] += float(r[ix["# Samples"]] or 0)
    thr[key] += float(r[ix["Thread Instructions Executed"]] or 0)
ti, ts = sum(inst.values()), sum(smp.values())
print(f"total warp instructions {ti:.4g}, samples {ts:.0f}, avg active threads {sum(thr.values())/ti:.1f}")
srcs = {}
for key, c in sorted(inst.items(), key=lambda kv: -(smp[kv[0]] if os.environ.get("BY_SMP") else kv[1]))[:top]:
    if key is None:
        text = "?"
    else:
        f = key[0]
        if f not in srcs:
            for root in ("integrated_path_planning_b200/csrc", "include", "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    srcs[f] = open(p, errors="ignore").read().splitlines()
                    break
            else:
                srcs[f] = []
        text = srcs[f][key[1] - 1].strip()[:100] if key[1] - 1 < len(srcs[f]) else ""
    print(f"{c/ti*100:5.1f}% inst {smp[key]/ts*100:5.1f}% smp  {key}: {text}")
