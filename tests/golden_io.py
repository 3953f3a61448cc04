"""Readers for the gzipped golden results and check.test.py's comparison."""
import gzip
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "ref_results")


def golden_path(case, name):
    return os.path.join(REF, case, name + ".gz")


def read_tokens(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        return f.read().split()


def read_text(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        return f.read()


def compare_txt_files(ref_path, test_path, threshold=1e-3):
    """check.test.py:40-52: token-wise abs(a-b) <= threshold and equal token counts."""
    a, b = read_tokens(ref_path), read_tokens(test_path)
    if len(a) != len(b):
        return False, "token count %d vs %d" % (len(a), len(b))
    fa, fb = np.array(a, dtype=np.float64), np.array(b, dtype=np.float64)
    d = np.abs(fa - fb)
    bad = int(np.sum(d > threshold))
    return bad == 0, "max abs diff %.3e, %d tokens over %g" % (float(d.max()) if d.size else 0.0, bad, threshold)


def load_frt(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        return np.loadtxt(f)
