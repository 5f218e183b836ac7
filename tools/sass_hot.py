"""Join an ncu SASS source page (ncu -i rep --page source --csv) with nvdisasm line info and rank source lines.

usage: python tools/sass_hot.py <report.source.csv> <lib.so> <cubin-name-in-lib> <mangled-kernel-substring> [top]
e.g.   python tools/sass_hot.py gpurun_out/x/fwd.source.csv deepdow_b200/libddw_b200.so markowitz_fwd_f32 fwd_sparse_kernelIfLi128
"""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, lib, cubin, kern = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(src_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[hdr_i])}
recs = rows[hdr_i + 1:]
tmp = tempfile.mkdtemp()
name = f"{cubin}.sm_100a.cubin"
subprocess.run(["cuobjdump", "-xelf", name, os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, name)], capture_output=True, text=True).stdout
lines, cur, infunc = [], ("?", 0), False
for ln in out.splitlines():
    if ln.startswith(".text."):
        infunc = kern in ln
        continue
    if ln.startswith("//---") or ln.lstrip().startswith(".section"):
        if ".text." in ln and kern not in ln:
            infunc = False
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if infunc and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
print(f"{len(recs)} SASS rows in report, {len(lines)} instructions with line info")
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
stall_cols = [n for n in col if n.startswith("stall_") and "Not Issued" not in n]
n = min(len(recs), len(lines))
tot_s = tot_i = 0
for r, l in zip(recs[:n], lines[:n]):
    s = int(r[col["# Samples"]] or 0); i = int(r[col["Instructions Executed"]] or 0)
    a = agg[l]; a[0] += s; a[1] += i; a[2] += 1
    for sc in stall_cols:
        v = int(r[col[sc]] or 0)
        if v: a[3][sc[6:]] += v
    tot_s += s; tot_i += i
print(f"total samples {tot_s}, warp instructions {tot_i}")
for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = " ".join(f"{k}:{v}" for k, v in a[3].most_common(3))
    print(f"{l[0]:24s}:{l[1]:5d} samp {a[0]:6d} ({100*a[0]/max(tot_s,1):4.1f}%) inst {a[1]:9d} ({100*a[1]/max(tot_i,1):4.1f}%) sass {a[2]:4d} | {st}")
