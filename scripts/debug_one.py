import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vcfdist_b200 import capi
from vcfdist_b200.batch import BatchBuilder, TYPE_SUB, TYPE_INS, TYPE_DEL
import ctypes as C
os.environ["VD_FORCE_CLASS"] = "1"
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libvcfdist_b200_dbg.so")
bb = BatchBuilder()
bb.add(b"ACGT", [[], [], [], []])
b = bb.build()
e = capi.Engine(0)
got = e.run(b).trimmed(); want = capi.oracle_run(b).trimmed()
for k in got: print(k, got[k], want[k])
