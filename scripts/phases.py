"""Per-phase cycle split of wsc_block_kernel (library built with -DVD_PHASE_PROF: scripts/build_lib.sh -DVD_PHASE_PROF -o <path>,
VD_LIB=<path>).  usage: phases.py [n_sc]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vcfdist_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_600_000
capi._lib = capi.load_library(os.environ["VD_LIB"])
b, cells, total = bench.make_workload("wgs", n, 1, 0, 1, 10000)
e = capi.Engine(0)
out = None
for i in range(3):
    out = e.run(b, out)
buf = (C.c_ulonglong * 56)()
capi._lib.vd_debug_phases(buf, 1)
out = e.run(b, out)
capi._lib.vd_debug_phases(buf, 0)
names = ["desc", "expand(P1)", "sweeps(P2)", "walk(P3)"]
for k in range(8):
    row = [buf[6 * k + i] for i in range(6)]
    if not row[4]:
        continue
    tot = sum(row[:4])
    print(f"S={k // 2 + 1} hom={k % 2}: blocks {row[4]}, cycles/block {tot / row[4]:.0f}: " +
          ", ".join(f"{nm} {100 * row[i] / tot:.1f}%" for i, nm in enumerate(names)))
for o, nm in ((48, "small kernels"), (52, "wsc / long")):
    print(f"walk_credit on thread 0 of a block ({nm}): walk loop {buf[o]} cycles, credit loop {buf[o + 1]} cycles, {buf[o + 2]} path steps, {buf[o + 3]} sync sections")
e.close()
