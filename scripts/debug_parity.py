import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden, OUT_KEYS
from oracle import checkers
from vcfdist_b200 import capi
from workloads import synth
name = sys.argv[1] if len(sys.argv) > 1 else "adv_11"
b, _, _ = load_golden(name)
for fc in (os.environ.get("DBG_CLASSES", "1").split(",")):
    if fc: os.environ["VD_FORCE_CLASS"] = fc
    e = capi.Engine(0)
    got = e.run(b).trimmed(); want = checkers.oracle_run(b).trimmed()
    L = b.hap_len(); lr = b.window_len()
    bad_aln = np.flatnonzero((got["aln_score"] != want["aln_score"]) | (got["status"] != want["status"]) | (got["aln_end_plane"] != want["aln_end_plane"])| (got["aln_beg_plane"] != want["aln_beg_plane"]))
    print("force", fc, "stats", e.stats())
    print("bad alignments", len(bad_aln), "of", 4*b.n_sc)
    for i in bad_aln[:12]:
        sc, ai = i // 4, i % 4
        print(f" sc {sc} ai {ai} lr {lr[sc]} L {L[sc]} got s={got['aln_score'][i]} e={got['aln_end_plane'][i]} b={got['aln_beg_plane'][i]} st={got['status'][i]:x} | want s={want['aln_score'][i]} e={want['aln_end_plane'][i]} b={want['aln_beg_plane'][i]} st={want['status'][i]:x}")
    for k in OUT_KEYS[4:]:
        print(k, int((got[k] != want[k]).sum()))
    e.close()
