#!/usr/bin/env python
"""Per-source-line attribution of an ncu capture: joins `ncu --page source --csv` (SASS rows with
executed-instruction counts and stall samples) with `nvdisasm -g` line info of the same cubin.
usage: ncu_lines.py <report.ncu-rep> <libvcfdist_b200.so> <kernel-substring> [top]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "finalize" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# offset -> (file, line) for the wanted kernel
line_of, cur, infn = {}, None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        infn = kname in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname.split("ILi")[0]], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
body = [r for r in rows[hi + 1:] if len(r) > 5 and r[0].startswith("0x")]
base = int(body[0][0], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in body:
    off = int(r[ci["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    v = [int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0)]
    for i in range(3):
        agg[key][i] += v[i]; tot[i] += v[i]
print(f"kernel {kname}: samples {tot[0]}, warp-instr {tot[1]}, thread-instr {tot[2]} (SIMT eff {tot[2]/max(1,tot[1])/32:.2f})")
src_cache = {}
def src(f, n):
    for d in ("vcfdist_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
    return ""
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/max(1,tot[0]):5.1f}% smp {100*v[1]/max(1,tot[1]):5.1f}% ins  {key[0]}:{key[1]:<4d} {src(*key)}")
