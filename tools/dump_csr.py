#!/usr/bin/env python
"""Writes the synthetic training set of a named shape as the binary CSR that oracle/_ref/shim-e2e reads
(int64 d1, d2, nnz; int64 row_ptr[d1+1]; int32 item[nnz]; float64 rating[nnz]) and, with --dir, as a reference data
directory (meta + training.ratings text) for the CLIs.  Measurement plumbing only."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="netflix")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--dir", default="")
    args = ap.parse_args()
    import torch
    from primalcr_b200.data import Dataset, Ratings, synth_dataset, write_reference_dir
    ds = synth_dataset(args.workload, scale=args.scale, device="cuda" if torch.cuda.is_available() else "cpu", test_per_user=0)
    R = ds.train
    if args.out:
        with open(args.out, "wb") as f:
            np.array([R.d1, R.d2, R.nnz], np.int64).tofile(f)
            R.row_ptr.astype(np.int64).tofile(f); R.item.astype(np.int32).tofile(f); R.rating.astype(np.float64).tofile(f)
    if args.dir:
        t = time.time()
        write_reference_dir(args.dir, Dataset(R, Ratings.empty(R.d1, R.d2)))
        print("wrote %s in %.1fs" % (args.dir, time.time() - t))
    print("d1=%d d2=%d nnz=%d" % (R.d1, R.d2, R.nnz))


if __name__ == "__main__":
    main()
