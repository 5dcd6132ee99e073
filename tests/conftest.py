import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """-> (Batch, refA dict, refB dict) from tests/golden/<name>.npz"""
    from vcfdist_b200.batch import Batch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    b = Batch(ref_off=z["ref_off"], ref_seq=z["ref_seq"], var_off=z["var_off"], var_pos=z["var_pos"],
              var_rlen=z["var_rlen"], var_type=z["var_type"], alt_off=z["alt_off"], alt_seq=z["alt_seq"],
              var_qual=z["var_qual"], max_qual=float(z["max_qual"]),
              rplane_seq=z["rplane_seq"] if "rplane_seq" in z.files else None)
    keys = ("errtypes", "sync_group", "ref_ed", "query_ed", "callq", "credit", "sc_phase", "orig_dist", "swap_dist")
    refA = {k: z["refA_" + k] for k in keys}
    refB = {k: z["refB_" + k] for k in keys}
    return b, refA, refB


FINAL_KEYS = ("errtypes", "sync_group", "ref_ed", "query_ed", "callq", "credit", "sc_phase", "orig_dist", "swap_dist")
OUT_KEYS = ("aln_score", "aln_end_plane", "aln_beg_plane", "status", "assigned", "sync_group", "ref_ed",
            "query_ed", "callq")


def mismatches(a: dict, b: dict, keys, var_mask=None):
    """Bit-exact comparison (NaN == NaN); returns {key: count} of differing elements."""
    bad = {}
    for k in keys:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, (k, x.shape, y.shape)
        if x.dtype.kind == "f":
            neq = x.view(np.uint32) != y.view(np.uint32)
            neq &= ~(np.isnan(x) & np.isnan(y))
        else:
            neq = x != y
        if var_mask is not None and k in FINAL_KEYS[:6]:
            neq = neq & var_mask
        if neq.any():
            bad[k] = int(neq.sum())
    return bad


def non_tie_var_mask(batch, status):
    """[2*n_var] True for variants of superclusters in which no alignment raised VD_ST_TIE."""
    tie_sc = (status[: 4 * batch.n_sc].reshape(-1, 4) & 1).any(axis=1)
    sc_of_var = np.repeat(np.arange(batch.n_sc), np.diff(batch.var_off[::4]))
    ok = ~tie_sc[sc_of_var]
    return np.concatenate([ok, ok]), int(tie_sc.sum())


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
