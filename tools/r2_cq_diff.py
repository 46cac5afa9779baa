"""Record-by-record comparison of the hit records two builds returned for the same ray batches (tools/trace_ab3.py with AB_DUMP).
python tools/r2_cq_diff.py /tmp/cqd/h /tmp/cqd/h_cq"""
import glob
import sys

import numpy as np

a_prefix, b_prefix = sys.argv[1], sys.argv[2]
for a in sorted(glob.glob(a_prefix + "_[0-9]*.npy")):
    b = b_prefix + a[len(a_prefix):]
    x, y = np.load(a), np.load(b)
    d = (x != y).reshape(len(x), -1).any(axis=1)
    print(a.split("/")[-1], "records", len(x), "differing", int(d.sum()), flush=True)
